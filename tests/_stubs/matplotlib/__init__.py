"""Empty stand-in so that `import utils` of the reference works where matplotlib is absent
(utils/util_func.py:6 imports matplotlib.pyplot; only showCurve uses it)."""
