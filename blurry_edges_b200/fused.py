"""Fused siblings of the script-level classes of the reference.

`PostProcessFused` has the constructor and `forward` signature of `blurry_edges_test.PostProcess`
(blurry_edges_test.py:12-100) but runs pass A / pass B as the fused sm_100a kernels of
csrc/be_kernels.cu through the C ABI.  A script switches with a one-line import."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


def want_deterministic(flag):
    """flag None: follow torch.use_deterministic_algorithms (what utils/util_func.py:17-19 sets for global_training.py:177)."""
    return torch.are_deterministic_algorithms_enabled() if flag is None else bool(flag)


def _geometry_from_args(args):
    H, W = int(args.img_size[0]), int(args.img_size[1])
    return dict(R=int(args.R), stride=int(args.stride), H=H, W=W, w=float(args.w), alpha_lambda=float(args.alpha_lambda),
                cam=dict(args.cam_params), mag=float(args.mag), rho_prime=float(getattr(args, 'rho_prime', 10.39)))


class PostProcessFused(nn.Module):
    """Drop-in for `blurry_edges_test.PostProcess(args, depthCal, device)`.

    forward(est, ny_pat, colors_only=True):
      colors_only=True   est [2B,L,10] (xy, wrapped angles, eta coefficients), ny_pat = the unfolded patches
                         [2B,3,R,R,Hp,Wp] of the reference call site (blurry_edges_test.py:120,128) or the images
                         themselves [2B,3,H,W]  ->  colours [2B,3,3,Hp,Wp] (device tensor)
      colors_only=False  est [B,L,12] restored params, ny_pat = unfolded [2,3,R,R,Hp,Wp] (B=1, as in the script),
                         [B,2,3,R,R,Hp,Wp], or images [B,2,3,H,W] / dataset-native [B,2,H,W,3]
                         ->  the six maps of blurry_edges_test.py:100 as NumPy arrays on the host
                         (`as_numpy=False` at construction keeps them as device tensors, like the big-image variant).
    """

    def __init__(self, args, depthCal=None, device='cuda:0', as_numpy=True, max_batch=None, deterministic=None):
        super().__init__()
        self.device = torch.device(device)
        self.depthCal = depthCal
        self.deterministic = deterministic          # None: follow torch.are_deterministic_algorithms_enabled()
        self.R, self.stride, self.w = int(args.R), int(args.stride), float(args.w)
        self.batch_size = int(args.batch_size)
        self.H, self.W = int(args.img_size[0]), int(args.img_size[1])
        self.rho_prime = float(getattr(args, 'rho_prime', 10.39))
        self.densify = getattr(args, 'densify', None)
        self.as_numpy = as_numpy
        self._geo = _geometry_from_args(args)
        self.ctx = _lib.Context(_lib.make_config(max_batch=int(max_batch or self.batch_size), **self._geo), self.device)
        self.H_patches, self.W_patches = self.ctx.Hp, self.ctx.Wp
        self.lambda_ridge = self.ctx.lambda_ridge
        self._num_patches = None
        self.last_depth_thresholded = None

    # PostProcessGlobalBase.num_patches (utils/postprocessing_loss.py:139-143)
    @property
    def num_patches(self):
        if self._num_patches is None:
            self._num_patches = self.ctx.cover_count()
        return self._num_patches

    def _grow(self, B):
        if B > self.ctx.max_batch:
            if torch.cuda.is_current_stream_capturing():      # a captured graph would keep pointers into the workspace freed here
                raise _lib.BlurryEdgesError(f'batch of {B} pairs exceeds the context ({self.ctx.max_batch}) inside a CUDA-graph capture: '
                                            'construct the helper with max_batch >= the largest batch')
            self.ctx.close()
            self.ctx = _lib.Context(_lib.make_config(max_batch=B, **self._geo), self.device)

    def _f32(self, t):
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    def _as_image(self, t, pair):
        """-> (tensor, layout).  Accepts unfolded patches or images (see class docstring)."""
        H, W, R = self.H, self.W, self.R
        t = self._f32(t)
        tail = (3, R, R, self.H_patches, self.W_patches)
        if tuple(t.shape[-5:]) == tail and t.dim() in (5, 6, 7):
            img = self.ctx.refold(t)                                   # [M,3,H,W]
            return img, (_lib.planar_layout(H, W) if pair else _lib.single_planar_layout(H, W))
        if pair:
            if t.dim() == 4:
                t = t.unsqueeze(0)
            if t.dim() == 5 and tuple(t.shape[1:]) == (2, 3, H, W):
                return t, _lib.planar_layout(H, W)
            if t.dim() == 5 and tuple(t.shape[1:]) == (2, H, W, 3):
                return t, _lib.channels_last_layout(H, W)
        elif t.dim() == 4 and tuple(t.shape[1:]) == (3, H, W):
            return t, _lib.single_planar_layout(H, W)
        raise _lib.BlurryEdgesError(f'cannot interpret image/patch tensor of shape {tuple(t.shape)} for a {H}x{W} image, R={R}')

    def forward(self, est, ny_pat, colors_only=True):
        est = self._f32(est)
        L = self.ctx.L
        if colors_only:
            if est.dim() != 3 or est.shape[1:] != (L, 10):
                raise _lib.BlurryEdgesError(f'colors_only expects est [2B,{L},10], got {tuple(est.shape)}')
            self._grow((est.shape[0] + 1) // 2)
            img, layout = self._as_image(ny_pat, pair=False)
            if img.shape[0] != est.shape[0]:
                raise _lib.BlurryEdgesError(f'{est.shape[0]} parameter sets but {img.shape[0]} images')
            return self.ctx.colors(est, img, layout, _lib.PARAMS_LOCAL10)
        if est.dim() != 3 or est.shape[1:] != (L, 12):
            raise _lib.BlurryEdgesError(f'expects est [B,{L},12], got {tuple(est.shape)}')
        B = est.shape[0]
        self._grow(B)
        img, layout = self._as_image(ny_pat, pair=True)
        if img.shape[0] * (1 if img.dim() == 5 else 0.5) != B:
            raise _lib.BlurryEdgesError(f'{B} parameter sets but image tensor {tuple(img.shape)}')
        self.ctx.set_deterministic(want_deterministic(self.deterministic))
        out = self.ctx.render_fold(est, img, layout, densify_w=(self.densify == 'w'), param_mode=_lib.PARAMS_RESTORED12)
        self.last_depth_thresholded = out[6]
        maps = out[:6]
        if self.as_numpy:
            return tuple(m.detach().cpu().numpy() for m in maps)       # blurry_edges_test.py:100
        return tuple(maps)


class PostProcessLocalFused(nn.Module):
    """Drop-in for `global_data_pre_cal.PostProcess(args, device)` (global_data_pre_cal.py:35-50): pass A on a flat batch of
    patches.  forward(params [N,10], pat_ny [N,R,R,3]) -> colours [N,3(channel),3(wedge)]."""

    def __init__(self, args, device='cuda:0'):
        super().__init__()
        self.device = torch.device(device)
        self.R, self.w, self.batch_size = int(args.R), float(args.w), int(args.batch_size)
        self._geo = dict(R=self.R, stride=1, H=self.R, W=self.R, w=self.w, alpha_lambda=float(args.alpha_lambda),
                         cam=dict(args.cam_params), mag=float(args.mag))
        self.ctx = None
        self.lambda_ridge = (float(args.alpha_lambda) * self.R ** 2) ** 2

    def get_colors(self, params, pat_ny):
        N, R = params.shape[0], self.R
        if params.dim() != 2 or params.shape[1] != 10 or tuple(pat_ny.shape) != (N, R, R, 3):
            raise _lib.BlurryEdgesError(f'expects params [N,10] and pat_ny [N,{R},{R},3], got {tuple(params.shape)}, {tuple(pat_ny.shape)}')
        if self.ctx is None or 2 * self.ctx.max_batch < N:
            if self.ctx is not None:
                self.ctx.close()
            self.ctx = _lib.Context(_lib.make_config(max_batch=(N + 1) // 2, **self._geo), self.device)
        est = params.to(device=self.device, dtype=torch.float32).reshape(N, 1, 10).contiguous()
        pat = pat_ny.to(device=self.device, dtype=torch.float32).contiguous()
        lay = _lib.BeImageLayout(3 * R * R, 0, 1, 3 * R, 3)
        return self.ctx.colors(est, pat, lay, _lib.PARAMS_LOCAL10).view(N, 3, 3)

    def forward(self, params, pat_ny):
        return self.get_colors(params, pat_ny)
