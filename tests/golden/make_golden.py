"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported read-only from
/root/reference) on the hash-synthesised inputs of tests/synth.py.

Run in the build container only:   python tests/golden/make_golden.py
Inputs are NOT stored (tests regenerate them bit-identically from tests/synth.py), except the
network-produced `est` of the config-1 case.  Every stored array names the reference call that
made it in the comment next to it."""
from __future__ import annotations

import math
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refimport  # noqa: E402
import synth  # noqa: E402

F64, F32 = torch.float64, torch.float32
GEOMS = {'tiny': 29, 'mid': 45}          # 5x5 and 13x13 patches (mid reaches the full 121-fold overlap)


def restore(raw):
    # blurry_edges_test.py:135-138 (script glue, restated)
    return torch.cat([raw[..., :4] * 3, torch.remainder((raw[..., 4:8] + 1) * math.pi, 2 * math.pi), raw[..., 8:] + 0.5], -1)


def unfold_pair(img_b2hw3, R, s):
    # blurry_edges_test.py:119-120
    t = img_b2hw3.flatten(0, 1).permute(0, 3, 1, 2)
    Hp = (t.shape[2] - R) // s + 1
    Wp = (t.shape[3] - R) // s + 1
    return torch.nn.Unfold(R, stride=s)(t).view(2, 3, R, R, Hp, Wp)


def np_(t):
    return t.detach().cpu().numpy()


def gen_inference(out):
    bt = refimport.module('blurry_edges_test')
    utils = refimport.module('utils')
    for gname, S in GEOMS.items():
        for densify in (None, 'w'):
            argv = ['--cuda', 'cpu', '--img_size', S, S, '--batch_size', 1] + (['--densify', densify] if densify else [])
            args = refimport.get_args('eval', argv)
            for dt, tag in ((F64, 'f64'), (F32, 'f32')):
                cal = utils.DepthEtas(args, 'cpu')
                pp = refimport.to_dtype(bt.PostProcess(args, cal, 'cpu'), cal, dt)
                L = pp.H_patches * pp.W_patches
                img = synth.image_pairs(1, S, S, seed=3, dtype=dt)
                pat = unfold_pair(img, args.R, args.stride)
                if densify is None:
                    estA = synth.est_local(2, L, seed=5, dtype=dt)
                    colors = pp(estA, pat, colors_only=True)                 # blurry_edges_test.py:81-92
                    out[f'{gname}/passA/{tag}'] = np_(colors)
                for kind in ('normal', 'stress'):
                    # parameters restored in fp32, as the script does (blurry_edges_test.py:135-138), then cast: the fp64 run of the
                    # reference sees bit-identical inputs to what the fp32 kernels are fed
                    est = restore(synth.raw_global(1, L, seed=7, kind=kind, dtype=F32)).to(dt)
                    maps = pp(est, pat, colors_only=False)                   # blurry_edges_test.py:81-100
                    key = f'{gname}/passB/{densify or "none"}/{kind}/{tag}'
                    for n, m in zip(('image', 'sharp', 'refoc', 'bndry', 'depth', 'conf'), maps):
                        out[f'{key}/{n}'] = m
    # global_data_pre_cal.PostProcess (local layout, N patches)              global_data_pre_cal.py:39-50
    pre = refimport.module('global_data_pre_cal')
    args = refimport.get_args('global_pre', ['--cuda', 'cpu'])
    for dt, tag in ((F64, 'f64'), (F32, 'f32')):
        pp = refimport.to_dtype(pre.PostProcess(args, 'cpu'), None, dt)
        est = synth.est_local(1, 40, seed=21, dtype=dt)[0]
        pat = synth.image_pairs(40, 21, 21, seed=22, dtype=dt)[:, 0]
        out[f'precal/{tag}'] = np_(pp(est, pat))


GAMMA_SETS = {
    'idx0': None,                                    # update_gamma() once -> gamma_idx 0
    'final': 'final',
    **{f'only{k}': k for k in range(7)},
}
GAMMA_ATTRS = ('gamma_color', 'gamma_color_cons', 'gamma_bndry_cons', 'gamma_smthns', 'gamma_smthns_cons',
               'gamma_bndry_loc', 'gamma_depth')


def gen_global_loss(out):
    gt_mod = refimport.module('global_training')
    utils = refimport.module('utils')
    B = 2
    for gname, S in GEOMS.items():
        args = refimport.get_args('global_train', ['--cuda', 'cpu', '--img_size', S, S, '--batch_size', B])
        for kind in ('normal', 'stress'):
            for dt, tag in ((F64, 'f64'), (F32, 'f32')):
                cal = utils.DepthEtas(args, 'cpu')
                crit = refimport.to_dtype(gt_mod.GlobalLoss(args, cal, 'cpu'), cal, dt)
                L = crit.H_patches * crit.W_patches
                img_ny = synth.image_pairs(B, S, S, seed=31, dtype=dt)
                img_gt, bd, deri, zgt = synth.loss_targets(B, S, S, seed=31, dtype=dt)
                for gname2, sel in GAMMA_SETS.items():
                    if tag == 'f32' and gname2 not in ('idx0', 'final'):
                        continue
                    if kind == 'stress' and gname2 not in ('idx0',):
                        continue
                    crit.gamma_idx = -1
                    crit.update_gamma()                                       # global_training.py:28-51
                    if sel == 'final':
                        crit.final_gamma()
                    elif isinstance(sel, int):
                        for k, a in enumerate(GAMMA_ATTRS):
                            setattr(crit, a, 1.0 if k == sel else 0.0)
                    raw = synth.raw_global(B, L, seed=33, kind=kind, dtype=dt).requires_grad_(True)
                    loss = crit(raw, img_ny, img_gt, bd, deri, zgt)           # global_training.py:147-157
                    (grad,) = torch.autograd.grad(loss, raw)
                    key = f'{gname}/gloss/{kind}/{gname2}/{tag}'
                    out[f'{key}/loss'] = np_(loss)
                    out[f'{key}/grad'] = np_(grad)
                    out[f'{key}/gammas'] = np.array([getattr(crit, a) for a in GAMMA_ATTRS])
                    if gname2 == 'idx0':
                        out[f'{key}/global_image'] = np_(crit.global_image)
                        out[f'{key}/global_bndry'] = np_(crit.global_bndry)


def gen_local_loss(out):
    lt = refimport.module('local_training')
    B = 8
    args = refimport.get_args('local_train', ['--cuda', 'cpu', '--batch_size', B])
    for dt, tag in ((F64, 'f64'), (F32, 'f32')):
        crit = refimport.to_dtype(lt.LocalLoss(args, 'cpu'), None, dt)
        est, ny, gt, bd, deri = synth.local_batch(B, args.R, seed=41, dtype=dt)
        for name, betas in (('final', None), ('loc', (1.0, 0.0)), ('smth', (0.0, 1.0))):
            crit.final_beta()                                                 # local_training.py:28-30
            if betas:
                crit.beta_bndry_loc, crit.beta_smthns = betas
            leaf = est.clone().requires_grad_(True)
            loss = crit(leaf * 1.0, ny, gt, bd, deri)                         # local_training.py:47-52
            (grad,) = torch.autograd.grad(loss, leaf)
            out[f'lloss/{name}/{tag}/loss'] = np_(loss)
            out[f'lloss/{name}/{tag}/grad'] = np_(grad)
            out[f'lloss/{name}/{tag}/betas'] = np.array([crit.beta_bndry_loc, crit.beta_smthns])


class _Capture:
    """Stand-in for utils.Visualizer: records the maps the unchanged driver hands over."""
    def __init__(self):
        self.calls = []

    def visualize(self, *a):
        self.calls.append([np.array(x) for x in a])
        return np.zeros((4, 4, 3), np.uint8)


def gen_config1(out):
    """Config 1: the unchanged depth_estimator (blurry_edges_test.py:102-172) with random-init nets."""
    bt = refimport.module('blurry_edges_test')
    utils = refimport.module('utils')
    models = refimport.module('models')
    S, alpha = 147, 190.0
    counts = synth.photon_pairs(1, S, S, seed=51, alpha=int(alpha))
    img = (torch.from_numpy(counts).float() / alpha)                          # data/dataset.py:63-73
    gt_depth = synth.uniform((1, S, S), 52, 0.75, 1.18)
    torch.manual_seed(0)
    local_m = models.LocalStage().eval()
    global_m = models.GlobalStage(in_parameter_size=38, out_parameter_size=12, device='cpu').eval()
    for densify in (None, 'w'):
        tmp = tempfile.mkdtemp()
        argv = ['--cuda', 'cpu', '--log_path', tmp] + (['--densify', densify] if densify else [])
        args = refimport.get_args('eval', argv)
        cal = utils.DepthEtas(args, 'cpu')
        rec = {}

        class Rec(bt.PostProcess):
            def forward(self, est, ny_pat, colors_only=True):
                if not colors_only:
                    rec['est'] = est.detach().clone()
                return super().forward(est, ny_pat, colors_only)

        helper = Rec(args, cal, 'cpu')
        cap = _Capture()
        printed = []
        import builtins
        old_print = builtins.print
        builtins.print = lambda *a, **k: printed.append(' '.join(str(x) for x in a))
        try:
            bt.depth_estimator(args, local_m, global_m, None, helper, cap, [(img, gt_depth)])
        finally:
            builtins.print = old_print
        I1, I2, C1, C2, Cs, Cr, conf, bnd, zgt, z = cap.calls[0]
        key = f'config1/{densify or "none"}'
        if densify is None:
            out['config1/est'] = np_(rec['est']).astype(np.float32)          # network-made: must be stored
        else:
            assert np.array_equal(out['config1/est'], np_(rec['est']).astype(np.float32))
        st = 6
        out[f'{key}/image'] = np.stack([C1, C2])[:, ::st, ::st]              # col_est[0,m] HxWx3, strided
        out[f'{key}/sharp'] = Cs[::st, ::st]
        out[f'{key}/refoc'] = Cr[::st, ::st]
        out[f'{key}/conf'] = conf[::st, ::st]
        out[f'{key}/bndry'] = bnd[::st, ::st]
        out[f'{key}/depth_thresholded'] = z[::st, ::st]
        m = utils.eval_depth(z[None], np_(gt_depth), z[None] > 0.0, crop=args.crop)  # utils/metrics.py:3-21
        out[f'{key}/metrics'] = np.array(m, dtype=np.float64)
        out[f'{key}/printed'] = np.array([l for l in printed if 'Error metrics' in l][0])


def gen_big(out):
    """Config 4 at 235x235 (2x2 blocks): unchanged blurry_edges_test_big.depth_estimator
    (blurry_edges_test_big.py:113-220) with stub networks that emit hash-synthesised params."""
    big = refimport.module('blurry_edges_test_big')
    utils = refimport.module('utils')
    S = 235
    tmp = tempfile.mkdtemp()
    args = refimport.get_args('eval', ['--cuda', 'cpu', '--log_path', tmp, '--big_img_size', S, S], big=True)
    cal = utils.DepthEtas(args, 'cpu')
    helper = big.PostProcess(args, cal, 'cpu')
    L = helper.H_patches * helper.W_patches
    calls = {'l': 0, 'g': 0}

    def local_stub(vec):
        k = calls['l']; calls['l'] += 1
        return synth.est_local(2, L, seed=60 + k).reshape(2 * L, 10)

    def global_stub(pm):
        k = calls['g']; calls['g'] += 1
        return synth.raw_global(1, L, seed=70 + k, kind='normal')

    img = torch.from_numpy(synth.photon_pairs(1, S, S, seed=61)).float() / 190.0
    gt_depth = synth.uniform((1, S, S), 62, 0.75, 1.18)
    cap = _Capture()
    import builtins
    old_print = builtins.print
    builtins.print = lambda *a, **k: None
    try:
        big.depth_estimator(args, local_stub, global_stub, helper, cap, [(img, gt_depth)])
    finally:
        builtins.print = old_print
    I1, I2, C1, C2, Cs, Cr, conf, bnd, zgt, z = cap.calls[0]
    st = 3
    out['big235/image'] = np.stack([C1, C2])[:, ::st, ::st]
    out['big235/sharp'] = Cs[::st, ::st]
    out['big235/refoc'] = Cr[::st, ::st]
    out['big235/conf'] = conf[::st, ::st]
    out['big235/bndry'] = bnd[::st, ::st]
    out['big235/depth_thresholded'] = z[::st, ::st]
    out['big235/nblocks'] = np.array(calls['g'])


def main():
    assert refimport.available(), 'reference not mounted'
    torch.set_num_threads(os.cpu_count())
    groups = {'inference': gen_inference, 'global_loss': gen_global_loss, 'local_loss': gen_local_loss,
              'config1': gen_config1, 'big': gen_big}
    only = sys.argv[1:] or list(groups)
    for name in only:
        out = {}
        groups[name](out)
        path = os.path.join(HERE, f'{name}.npz')
        np.savez_compressed(path, **{k.replace('/', '.'): v for k, v in out.items()})
        print(f'{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
    main()
