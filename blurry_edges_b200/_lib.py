"""ctypes binding of libblurry_edges_b200.so (C ABI in include/blurry_edges_b200.h).

There is deliberately no fallback: if the shared object is missing or a call fails, an exception
is raised.  PyTorch is used only for device memory and streams; tensors cross the boundary as raw
device pointers."""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libblurry_edges_b200.so')

PARAMS_RESTORED12, PARAMS_RAW12, PARAMS_LOCAL10, PARAMS_LOCALRAW10 = 0, 1, 2, 3


class BeConfig(C.Structure):
    _fields_ = [('R', C.c_int32), ('stride', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('w', C.c_double), ('alpha_lambda', C.c_double),
                ('cam_s', C.c_double), ('cam_rho_1', C.c_double), ('cam_rho_2', C.c_double),
                ('cam_sigma_cam', C.c_double), ('cam_pixel_pitch', C.c_double), ('cam_mag', C.c_double),
                ('rho_prime', C.c_double), ('max_batch', C.c_int32)]


class BeImageLayout(C.Structure):
    _fields_ = [('sb', C.c_int64), ('sm', C.c_int64), ('sc', C.c_int64), ('sy', C.c_int64), ('sx', C.c_int64)]


class BeBlock(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('img', 'oy', 'ox', 'py0', 'py1', 'px0', 'px1')]


class BlurryEdgesError(RuntimeError):
    pass


_P = C.c_void_p
_SIGS = {
    'be_abi_version': (C.c_int, []),
    'be_last_error': (C.c_char_p, []),
    'be_launch_count': (C.c_int64, []),
    'be_ctx_create': (C.c_int, [C.POINTER(_P), C.POINTER(BeConfig)]),
    'be_ctx_destroy': (C.c_int, [_P]),
    'be_ctx_workspace_bytes': (C.c_int64, [_P]),
    'be_ctx_constants': (C.c_int, [_P, C.POINTER(C.c_double)]),
    'be_derive_constants': (C.c_int, [C.POINTER(BeConfig), C.POINTER(C.c_double)]),
    'be_ctx_set_timing': (C.c_int, [_P, C.c_int32]),
    'be_ctx_set_deterministic': (C.c_int, [_P, C.c_int32]),
    'be_ctx_last_timing': (C.c_int, [_P, C.POINTER(C.c_float)]),
    'be_cover_count': (C.c_int, [_P, _P, _P]),
    'be_refold_image': (C.c_int, [_P, _P, C.c_int32, _P, _P]),
    'be_colors_fwd': (C.c_int, [_P, _P, C.c_int32, _P, C.POINTER(BeImageLayout), C.c_int32, _P, _P]),
    'be_render_fold_fwd': (C.c_int, [_P, _P, C.c_int32, _P, C.POINTER(BeImageLayout), C.c_int32, C.c_int32,
                                     _P, _P, _P, _P, _P, _P, _P, _P]),
    'be_colors_blocks_fwd': (C.c_int, [_P, _P, C.c_int32, _P, C.POINTER(BeImageLayout), C.POINTER(BeBlock), C.c_int32, _P, _P]),
    'be_render_fold_blocks': (C.c_int, [_P, _P, C.c_int32, _P, C.POINTER(BeImageLayout), C.POINTER(BeBlock), C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    'be_fold_normalise': (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_double, _P, _P, _P, _P, _P, _P, _P, _P]),
    'be_fold_normalise_band': (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, _P, _P, _P, _P, _P, _P, _P, _P]),
    'be_params2dists': (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int64, _P, _P]),
    'be_params2dists_bwd': (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, C.c_int64, _P, _P]),
    'be_dists2indicators': (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, _P, _P]),
    'be_dists2indicators_bwd': (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int64, _P, _P, _P]),
    'be_elementwise': (C.c_int, [_P, C.c_int32, _P, C.c_double, C.c_int64, _P, _P]),
    'be_elementwise_bwd': (C.c_int, [_P, C.c_int32, _P, _P, C.c_double, C.c_int64, _P, _P]),
    'be_smish': (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    'be_smish_bwd': (C.c_int, [_P, _P, _P, C.c_int64, _P, _P]),
    'be_etas2depth': (C.c_int, [_P, _P, _P, C.c_int64, _P, _P]),
    'be_etas2depth_bwd': (C.c_int, [_P, _P, _P, _P, C.c_int64, _P, _P, _P]),
    'be_inverse_3by3': (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    'be_inverse_3by3_bwd': (C.c_int, [_P, _P, _P, C.c_int64, _P, _P]),
    'be_image_derivative': (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P]),
    'be_image_derivative_bwd': (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P]),
    'be_fold': (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, _P]),
    'be_fold_depth': (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P]),
    'be_unfold': (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, _P]),
    'be_patch_gather': (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    'be_assemble_pm': (C.c_int, [_P, _P, _P, C.c_int64, _P, _P]),
    'be_eval_depth': (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    'be_global_loss_stage1': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, _P]),
    'be_global_loss_stage1_render': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P]),
    'be_global_loss_stage1_targets': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, _P]),
    'be_global_loss_stage2': (C.c_int, [_P, C.c_int32, C.POINTER(C.c_double), C.c_int64, _P, _P, _P, _P, _P]),
    'be_global_loss_stage2_launch': (C.c_int, [_P, C.c_int32, C.POINTER(C.c_double), C.c_int64, _P, _P, _P]),
    'be_global_loss_stage2_finish': (C.c_int, [_P, C.c_int32, C.POINTER(C.c_double), C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    'be_host_global_loss': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.POINTER(C.c_double), _P, _P, _P]),
    'be_host_global_loss_begin': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.POINTER(C.c_double), C.c_int64, C.c_int32, _P, _P]),
    'be_host_global_loss_end': (C.c_int, [_P, C.c_int32, C.POINTER(C.c_double), C.c_int64, _P, _P, _P, _P, _P, _P]),
    'be_ctx_last_train_timing': (C.c_int, [_P, C.POINTER(C.c_float)]),
    'be_ctx_train_timing_at': (C.c_int, [_P, C.c_int32, C.POINTER(C.c_float)]),
    'be_local_loss': (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_double, C.c_double, C.c_int32, _P, _P, _P, _P]),
    'be_host_render_fold': (C.c_int, [_P, _P, C.c_int32, _P, C.POINTER(BeImageLayout), C.c_int32, C.c_int32,
                                      _P, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def exported_symbols():
    """Names declared in include/blurry_edges_b200.h that the library must export."""
    return list(_SIGS)


def load():
    """dlopen the library (building is the job of blurry_edges_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BlurryEdgesError(f'{LIB_PATH} is missing: run `python -m blurry_edges_b200.build` '
                               '(there is no CPU or PyTorch fallback for this path)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise BlurryEdgesError(load().be_last_error().decode())


def launch_count() -> int:
    return int(load().be_launch_count())


def make_config(R=21, stride=2, H=147, W=147, w=1.0, alpha_lambda=5e-3, cam=None, mag=4.0, rho_prime=10.39,
                max_batch=1) -> BeConfig:
    cam = cam or {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}
    return BeConfig(R, stride, H, W, w, alpha_lambda, cam['s'], cam['rho_1'], cam['rho_2'], cam['sigma_cam'],
                    cam['pixel_pitch'], mag, rho_prime, max_batch)


def derive_constants(cfg: BeConfig):
    out = (C.c_double * 8)()
    check(load().be_derive_constants(C.byref(cfg), out))
    return list(out)


def _ptr(t: torch.Tensor | None):
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise BlurryEdgesError(f'expected a contiguous float32 CUDA tensor, got {t.dtype} on {t.device}, '
                               f'contiguous={t.is_contiguous()}')
    return C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def planar_layout(H, W) -> BeImageLayout:
    """[B,2,3,H,W] (or [M,3,H,W] with sm unused)."""
    return BeImageLayout(6 * H * W, 3 * H * W, H * W, W, 1)


def single_planar_layout(H, W) -> BeImageLayout:
    """[M,3,H,W]: one image per batch index."""
    return BeImageLayout(3 * H * W, 0, H * W, W, 1)


def channels_last_layout(H, W) -> BeImageLayout:
    """dataset-native [B,2,H,W,3] (data/dataset.py:63)."""
    return BeImageLayout(6 * H * W, 3 * H * W, 1, 3 * W, 3)


class Context:
    """Owns one be_ctx (geometry + camera + HBM workspace) on one device."""

    def __init__(self, cfg: BeConfig, device):
        self.lib = load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise BlurryEdgesError(f'blurry_edges_b200 runs on CUDA devices only (got {self.device}); no CPU fallback')
        self.cfg = cfg
        self.h = _P()
        with torch.cuda.device(self.device):
            check(self.lib.be_ctx_create(C.byref(self.h), C.byref(cfg)))
        c = (C.c_double * 8)()
        check(self.lib.be_ctx_constants(self.h, c))
        (self.numerator, self.denominator_constant, self.denominator_factor_root, self.denominator_factor,
         self.intercept, self.lambda_ridge) = list(c)[:6]
        self.Hp, self.Wp = int(c[6]), int(c[7])
        self.L = self.Hp * self.Wp

    def close(self):
        if getattr(self, 'h', None) and self.h.value:
            self.lib.be_ctx_destroy(self.h)
            self.h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def max_batch(self):
        return self.cfg.max_batch

    def workspace_bytes(self):
        return int(self.lib.be_ctx_workspace_bytes(self.h))

    def set_deterministic(self, enable=True):
        """Fixed-order fold (bit-identical results from launch to launch) instead of the atomic fold; see the header."""
        enable = bool(enable)
        if getattr(self, '_det', False) != enable:
            with torch.cuda.device(self.device):
                check(self.lib.be_ctx_set_deterministic(self.h, int(enable)))
            self._det = enable

    def set_timing(self, enable=True):
        with torch.cuda.device(self.device):
            check(self.lib.be_ctx_set_timing(self.h, int(enable)))

    def last_timing(self):
        """ms of (memset, setup kernel, run kernel, normalise kernel) of the last render_fold call."""
        ms = (C.c_float * 4)()
        with torch.cuda.device(self.device):
            check(self.lib.be_ctx_last_timing(self.h, ms))
        return list(ms)

    def last_train_timing(self, steps_back=0):
        """ms of (memset, setup, be_run3_kernel<TRAINFWD>, ~0 [slot of the former normalise launch], be_train_targets_kernel, be_loss2_kernel,
        reduce + depth fix-up) of
        the last global-loss step, or of the step `steps_back` (< 64) before it (waits for that step)."""
        ms = (C.c_float * 7)()
        with torch.cuda.device(self.device):
            check(self.lib.be_ctx_train_timing_at(self.h, int(steps_back), ms))
        return list(ms)

    # ---- device-pointer entry points ---------------------------------------------------
    def cover_count(self):
        out = torch.empty(self.cfg.H, self.cfg.W, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.be_cover_count(self.h, _ptr(out), _stream(self.device)))
        return out

    def refold(self, unfolded: torch.Tensor):
        M = unfolded.numel() // (3 * self.cfg.R * self.cfg.R * self.L)
        img = torch.empty(M, 3, self.cfg.H, self.cfg.W, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.be_refold_image(self.h, _ptr(unfolded), M, _ptr(img), _stream(self.device)))
        return img

    def colors(self, est: torch.Tensor, img: torch.Tensor, layout: BeImageLayout, param_mode=PARAMS_LOCAL10):
        M = est.shape[0]
        out = torch.empty(M, 3, 3, self.Hp, self.Wp, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.be_colors_fwd(self.h, _ptr(est), param_mode, _ptr(img), C.byref(layout), M, _ptr(out),
                                         _stream(self.device)))
        return out

    def render_fold(self, est: torch.Tensor, img: torch.Tensor, layout: BeImageLayout, densify_w=False,
                    param_mode=PARAMS_RESTORED12, out=None, want_thresholded=True):
        B, H, W = est.shape[0], self.cfg.H, self.cfg.W
        if out is None:
            out = self.alloc_outputs(B, want_thresholded)
        with torch.cuda.device(self.device):
            check(self.lib.be_render_fold_fwd(self.h, _ptr(est), param_mode, _ptr(img), C.byref(layout), B, int(densify_w),
                                              *[_ptr(t) for t in out], *([None] if len(out) == 6 else []),
                                              _stream(self.device)))
        return out

    def alloc_outputs(self, B, want_thresholded=True):
        H, W, kw = self.cfg.H, self.cfg.W, dict(device=self.device, dtype=torch.float32)
        out = [torch.empty(B, 2, 3, H, W, **kw), torch.empty(B, 3, H, W, **kw), torch.empty(B, 3, H, W, **kw),
               torch.empty(B, 1, H, W, **kw), torch.empty(B, H, W, **kw), torch.empty(B, H, W, **kw)]
        if want_thresholded:
            out.append(torch.empty(B, H, W, **kw))
        return out

    def call(self, name, *args):
        """Generic launcher for the method-granularity entry points: tensors become device pointers, the current
        stream is appended."""
        for a in args:
            if isinstance(a, torch.Tensor) and not (a.is_cuda and a.is_contiguous()):
                raise BlurryEdgesError('expected contiguous CUDA tensors')
        conv = [(_ptr(a) if a.dtype == torch.float32 else C.c_void_p(a.data_ptr())) if isinstance(a, torch.Tensor) else a for a in args]
        with torch.cuda.device(self.device):
            check(getattr(self.lib, name)(self.h, *conv, _stream(self.device)))

    # ---- blocked (big-image) entry points --------------------------------------------------
    @staticmethod
    def _blocks(blocks):
        arr = (BeBlock * len(blocks))()
        for k, b in enumerate(blocks):
            arr[k] = BeBlock(*[int(v) for v in b])
        return arr

    def colors_blocks(self, est, img, layout, blocks, param_mode=PARAMS_LOCAL10):
        """est [nitem,L,10]; blocks = [(img, oy, ox, py0, py1, px0, px1)] per item -> colours [nitem,3,3,Hp,Wp]"""
        n = est.shape[0]
        out = torch.zeros(n, 3, 3, self.Hp, self.Wp, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.be_colors_blocks_fwd(self.h, _ptr(est), param_mode, _ptr(img), C.byref(layout), self._blocks(blocks), n,
                                                _ptr(out), _stream(self.device)))
        return out

    def render_fold_blocks(self, est, img, layout, blocks, acc, densify_w=False, param_mode=PARAMS_RESTORED12, acc_y0=0):
        """adds the blocks' pass-B sums into acc [*,rows,accW,16] (caller-zeroed), which holds image rows [acc_y0, acc_y0 + rows)"""
        with torch.cuda.device(self.device):
            check(self.lib.be_render_fold_blocks(self.h, _ptr(est), param_mode, _ptr(img), C.byref(layout), self._blocks(blocks),
                                                 est.shape[0], int(densify_w), int(acc_y0), acc.shape[-3], acc.shape[-2], _ptr(acc),
                                                 _stream(self.device)))
        return acc

    def fold_normalise(self, acc, thres, y0=0, full_H=None, packed=False):
        """acc [B,rows,accW,16] holding image rows [y0, y0 + rows) of full_H (default: the whole image) ->
        (image, sharp, refoc, bndry, depth, conf, depth_thresholded) for those rows.  packed=True (B = 1): the seven maps are
        consecutive plane ranges of ONE [16, rows, W] tensor, returned as an eighth element (one message per rank in the band gather)."""
        B, H, W = acc.shape[0], acc.shape[1], acc.shape[2]
        full_H = H if full_H is None else int(full_H)
        kw = dict(device=self.device, dtype=torch.float32)
        if packed:
            if B != 1:
                raise BlurryEdgesError('packed outputs need a single image')
            buf = torch.empty(16, H, W, **kw)
            out = [buf[0:6].view(1, 2, 3, H, W), buf[6:9].view(1, 3, H, W), buf[9:12].view(1, 3, H, W), buf[12:13].view(1, 1, H, W),
                   buf[13:14].view(1, H, W), buf[14:15].view(1, H, W), buf[15:16].view(1, H, W)]
        else:
            out = [torch.empty(B, 2, 3, H, W, **kw), torch.empty(B, 3, H, W, **kw), torch.empty(B, 3, H, W, **kw),
                   torch.empty(B, 1, H, W, **kw), torch.empty(B, H, W, **kw), torch.empty(B, H, W, **kw), torch.empty(B, H, W, **kw)]
        if H > 0:
            with torch.cuda.device(self.device):
                check(self.lib.be_fold_normalise_band(self.h, _ptr(acc), B, int(y0), H, full_H, W, float(thres), *[_ptr(t) for t in out],
                                                      _stream(self.device)))
        return out + [buf] if packed else out

    # ---- training entry points -----------------------------------------------------------
    def global_loss_stage1(self, raw, img_ny, img_gt, bndry_dist, deri, bndry_depth, want_maps=True, between=None):
        """-> (global_image [B,2,3,H,W] | None, global_bndry [B,1,H,W] | None, counts int64[2] = (mask count, B*L): the pair a
        data-parallel caller all-reduces in one 16-byte collective).  `between(counts)`: called after the render kernels (the count
        is final once they finish) and before the target kernels are launched - where a data-parallel caller starts its all-reduce;
        its return value is handed back as a fourth element."""
        B, H, W, kw = raw.shape[0], self.cfg.H, self.cfg.W, dict(device=self.device, dtype=torch.float32)
        gimg = torch.empty(B, 2, 3, H, W, **kw) if want_maps else None
        gbnd = torch.empty(B, 1, H, W, **kw) if want_maps else None
        tmpl = getattr(self, '_cnt_template', None)
        if tmpl is None or tmpl[0] != B:
            tmpl = self._cnt_template = (B, torch.tensor([0, B * self.L], dtype=torch.int64).to(self.device))
        cnt = tmpl[1].clone()
        ptrs = (_ptr(raw), _ptr(img_ny), _ptr(img_gt), _ptr(bndry_dist), _ptr(deri), _ptr(bndry_depth))
        # the C side compares the two device pointers the same way (be_global_loss_stage1): the training call of the reference
        # passes one tensor twice (global_training.py:210) and gets the kernel variant that skips the GT target planes
        self.last_same_gt = img_ny.data_ptr() == img_gt.data_ptr()
        with torch.cuda.device(self.device):
            if between is None:
                check(self.lib.be_global_loss_stage1(self.h, *ptrs, B, _ptr(gimg), _ptr(gbnd), C.c_void_p(cnt.data_ptr()), _stream(self.device)))
                return gimg, gbnd, cnt
            check(self.lib.be_global_loss_stage1_render(self.h, *ptrs, B, C.c_void_p(cnt.data_ptr()), _stream(self.device)))
            mid = between(cnt)
            check(self.lib.be_global_loss_stage1_targets(self.h, *ptrs, B, _ptr(gimg), _ptr(gbnd), C.c_void_p(cnt.data_ptr()), _stream(self.device)))
        return gimg, gbnd, cnt, mid

    def global_loss_stage2(self, B, gammas, global_patches, mask_count, want_grad=True):
        """-> (terms [7], loss [1], grad [B,L,12] | None)"""
        kw = dict(device=self.device, dtype=torch.float32)
        terms, loss = torch.empty(7, **kw), torch.empty(1, **kw)
        grad = torch.empty(B, self.L, 12, **kw) if want_grad else None
        gam = (C.c_double * 7)(*[float(x) for x in gammas])
        with torch.cuda.device(self.device):
            check(self.lib.be_global_loss_stage2(self.h, B, gam, int(global_patches), C.c_void_p(mask_count.data_ptr()),
                                                 _ptr(terms), _ptr(loss), _ptr(grad), _stream(self.device)))
        return terms, loss, grad

    def global_loss_stage2_launch(self, B, gammas, global_patches, want_grad=True):
        """Start the loss kernel before the batch mask count is known (data-parallel: the count is being all-reduced).
        -> (grad [B,L,12] | None, grad_depth [B,L,4] | None): hand both to global_loss_stage2_finish."""
        kw = dict(device=self.device, dtype=torch.float32)
        grad = torch.empty(B, self.L, 12, **kw) if want_grad else None
        gdep = torch.empty(B, self.L, 4, **kw) if want_grad else None
        gam = (C.c_double * 7)(*[float(x) for x in gammas])
        with torch.cuda.device(self.device):
            check(self.lib.be_global_loss_stage2_launch(self.h, B, gam, int(global_patches), _ptr(grad), _ptr(gdep), _stream(self.device)))
        return grad, gdep

    def global_loss_stage2_finish(self, B, gammas, global_patches, counts, grad, grad_depth):
        """-> (terms [7], loss [1], grad): terms and loss from the kernel's partial sums and the (all-reduced) counts int64[2] =
        (mask count, true patch count of the global batch); the depth term's share of the gradient is normalised by the mask count
        and added to `grad` in place; if the true patch count differs from `global_patches` (uneven shards) everything is rescaled."""
        kw = dict(device=self.device, dtype=torch.float32)
        terms, loss = torch.empty(7, **kw), torch.empty(1, **kw)
        gam = (C.c_double * 7)(*[float(x) for x in gammas])
        true_p = C.c_void_p(counts.data_ptr() + 8) if counts.numel() > 1 else None
        with torch.cuda.device(self.device):
            check(self.lib.be_global_loss_stage2_finish(self.h, B, gam, int(global_patches), C.c_void_p(counts.data_ptr()), true_p,
                                                        _ptr(terms), _ptr(loss), _ptr(grad), _ptr(grad_depth), _stream(self.device)))
        return terms, loss, grad

    def local_loss(self, est, img_ny, img_gt, bndry_dist, deri, beta_bndry_loc, beta_smthns, want_grad=True, wrap_in_place=False):
        """-> (terms [3], loss [1], grad [B,10] | None); one kernel launch.  wrap_in_place: the angles est[:, 4:8] are written back
        wrapped to [0, 2 pi), as the reference does to the network output (local_training.py:33)."""
        B, kw = est.shape[0], dict(device=self.device, dtype=torch.float32)
        out = torch.empty(4, **kw)
        terms, loss = out[:3], out[3:]
        grad = torch.empty(B, 10, **kw) if want_grad else None
        with torch.cuda.device(self.device):
            check(self.lib.be_local_loss(self.h, _ptr(est), _ptr(img_ny), _ptr(img_gt), _ptr(bndry_dist), _ptr(deri), B,
                                         float(beta_bndry_loc), float(beta_smthns), int(wrap_in_place), C.c_void_p(terms.data_ptr()),
                                         C.c_void_p(loss.data_ptr()), _ptr(grad), _stream(self.device)))
        return terms, loss, grad

    def host_global_loss(self, raw, img_ny, img_gt, bndry_dist, deri, bndry_depth, gammas, want_grad=True, out=None, process_group=None,
                         deferred=False):
        """The training step on HOST tensors (CPU float32 contiguous, ideally pinned; dataset layouts of global_loss_stage1):
        -> (terms [7], loss [1], grad [B,L,12] | None) as CPU tensors (`out` = the same triple, reused).  With a process_group the
        mask count and the patch count of the global batch are all-reduced between the two halves of the call
        (be_host_global_loss_begin / _end: the depth normaliser is deferred); without one the two-phase schedule of be_host_global_loss
        runs, which knows the count before the loss kernels start.  deferred=True forces the begin / end pair on a single process."""
        B = raw.shape[0]
        ts = (raw, img_ny, img_gt, bndry_dist, deri, bndry_depth)
        for t in ts:
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise BlurryEdgesError('host_global_loss expects contiguous float32 CPU tensors')
        if out is None:
            pin = dict(dtype=torch.float32, pin_memory=True)
            out = (torch.empty(7, **pin), torch.empty(1, **pin), torch.empty(B, self.L, 12, **pin) if want_grad else None)
        terms, loss, grad = out
        gam = (C.c_double * 7)(*[float(x) for x in gammas])
        hp = [C.c_void_p(t.data_ptr()) for t in ts]
        gp = C.c_void_p(grad.data_ptr()) if grad is not None else None
        with torch.cuda.device(self.device):
            if process_group is None and not deferred:
                check(self.lib.be_host_global_loss(self.h, *hp, B, gam, C.c_void_p(terms.data_ptr()), C.c_void_p(loss.data_ptr()), gp))
            else:
                import torch.distributed as dist
                if getattr(self, '_host_cnt', None) is None:
                    self._host_cnt = torch.zeros(2, device=self.device, dtype=torch.int64)
                    self._host_cnt_B = torch.zeros(1, device=self.device, dtype=torch.int64)
                cnt = self._host_cnt
                # ONE 16-byte all-reduce of (mask count, patch count) between the two halves; the kernels are launched with the patch
                # count the host can know (own patches x ranks) and `end` rescales if the shards turn out to be uneven
                assumed = B * self.L * (dist.get_world_size(process_group) if process_group is not None else 1)
                check(self.lib.be_host_global_loss_begin(self.h, *hp, B, gam, assumed, int(grad is not None), C.c_void_p(cnt.data_ptr()),
                                                         _stream(self.device)))
                cnt[1:].fill_(B * self.L)
                if process_group is not None:
                    dist.all_reduce(cnt, group=process_group)
                check(self.lib.be_host_global_loss_end(self.h, B, gam, assumed, C.c_void_p(cnt.data_ptr()), C.c_void_p(cnt.data_ptr() + 8),
                                                       C.c_void_p(terms.data_ptr()), C.c_void_p(loss.data_ptr()), gp, _stream(self.device)))
        return terms, loss, grad

    # ---- host-buffer entry point (numpy / pinned host tensors in, numpy out) ------------
    def host_render_fold(self, est, img, layout: BeImageLayout, densify_w=False, param_mode=PARAMS_RESTORED12, out=None,
                         want_thresholded=True):
        """est, img: CPU float32 contiguous torch tensors (ideally pinned).  Returns the six maps of PostProcess.forward
        (blurry_edges_test.py:100) as CPU tensors, plus the thresholded depth of :144 if want_thresholded."""
        B, H, W = est.shape[0], self.cfg.H, self.cfg.W
        for t in (est, img):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise BlurryEdgesError('host_render_fold expects contiguous float32 CPU tensors')
        if out is None:
            pin = dict(dtype=torch.float32, pin_memory=True)
            out = [torch.empty(B, 2, 3, H, W, **pin), torch.empty(B, 3, H, W, **pin), torch.empty(B, 3, H, W, **pin),
                   torch.empty(B, 1, H, W, **pin), torch.empty(B, H, W, **pin), torch.empty(B, H, W, **pin)]
            if want_thresholded:
                out.append(torch.empty(B, H, W, **pin))
        ptrs = [C.c_void_p(t.data_ptr()) for t in out] + ([None] if len(out) == 6 else [])
        with torch.cuda.device(self.device):
            check(self.lib.be_host_render_fold(self.h, C.c_void_p(est.data_ptr()), param_mode, C.c_void_p(img.data_ptr()),
                                               C.byref(layout), B, int(densify_w), *ptrs))
        return out
