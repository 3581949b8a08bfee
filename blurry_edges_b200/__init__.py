"""blurry_edges_b200 - B200-native (sm_100a) render -> fold -> depth path of Blurry-Edges.

Only what the hot path needs lives here: csrc/ (CUDA kernels + C ABI), the ctypes binding (_lib) and the
host-side mirror of the reference's helper classes.  Importing the package never touches a GPU; constructing
any of its classes on a machine without the built library or without a CUDA device raises."""
from . import _lib
from ._lib import BlurryEdgesError, Context, make_config
from .base import DepthEtas, PostProcessBase, PostProcessGlobalBase, PostProcessLocalBase
from .fused import PostProcessFused, PostProcessLocalFused
from .pipeline import DepthEstimatorFused
from .losses import GlobalLossFused, LocalLossFused
from .big import BigImageFused, block_windows, shard_blocks
from .activations import SmishFused, patch_reference_smish, smish
from .data import ShapePrefetcher

__all__ = ['DepthEtas', 'PostProcessBase', 'PostProcessGlobalBase', 'PostProcessLocalBase', 'BlurryEdgesError', 'Context', 'make_config', 'PostProcessFused', 'PostProcessLocalFused', 'GlobalLossFused', 'LocalLossFused', 'BigImageFused', 'DepthEstimatorFused', 'block_windows', 'shard_blocks', 'SmishFused', 'patch_reference_smish', 'smish', 'ShapePrefetcher', '_lib']
