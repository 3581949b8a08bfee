"""Fused siblings of the training criteria of the reference:

  GlobalLossFused  ==  global_training.GlobalLoss   (global_training.py:11-157)
  LocalLossFused   ==  local_training.LocalLoss     (local_training.py:10-52)

Same constructor arguments, same gamma / beta schedules, same `forward` signatures; the loss and its gradient w.r.t. the
network output are produced by the sm_100a kernels of csrc/be_train.cu (forward + analytic backward in one pass) and
handed to autograd through a tiny torch.autograd.Function."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .fused import _geometry_from_args, want_deterministic


class _FusedLossFn(torch.autograd.Function):
    """loss = scale * f(est) with the gradient of f already computed by the kernel; backward only scales it."""

    @staticmethod
    def forward(ctx, est, loss, grad, scale=1.0):
        ctx.save_for_backward(grad)
        ctx.scale = scale
        return loss.reshape(()) if scale == 1.0 else loss.reshape(()) * scale

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * (gout if ctx.scale == 1.0 else gout * ctx.scale), None, None, None


class GlobalLossFused(nn.Module):
    """Data parallel (`process_group`): every rank holds a slice of the global batch and computes  local sums / normalisers of the
    WHOLE batch,  so the ranks' values ADD UP to the loss the reference computes on the whole batch in one process.
      grad_reduce='mean' (default): the returned loss (and its gradient) is multiplied by the number of ranks.  This is the
          setting for stock DistributedDataParallel, which AVERAGES parameter gradients over the ranks: the averaged gradient is then
          exactly the reference's whole-batch gradient (so clip_grad_norm_ and AdamW see the reference's step), and the mean over
          ranks of the returned losses is the reference's loss.
      grad_reduce='sum': the rank's share is returned unscaled; the caller sums losses / gradients over the ranks itself.
    `last_loss_share` always holds the unscaled share."""

    def __init__(self, args, depthCal=None, device='cuda:0', process_group=None, grad_reduce='mean', deterministic=None,
                 overlap_collective=True):
        super().__init__()
        self.overlap_collective = bool(overlap_collective)
        self.deterministic = deterministic          # None: follow torch.are_deterministic_algorithms_enabled() (global_training.py:177)
        if grad_reduce not in ('mean', 'sum'):
            raise _lib.BlurryEdgesError(f"grad_reduce must be 'mean' or 'sum', got {grad_reduce!r}")
        self.grad_reduce = grad_reduce
        self.device = torch.device(device)
        self.depthCal = depthCal
        self.R, self.stride, self.w = int(args.R), int(args.stride), float(args.w)
        self.batch_size = int(args.batch_size)
        self.H, self.W = int(args.img_size[0]), int(args.img_size[1])
        # global_training.py:14-22
        self.dynamic_epoch = args.dynamic_epoch
        self.gamma_color_range = args.gamma_color
        self.gamma_color_cons_range = args.gamma_color_cons
        self.gamma_bndry_cons_range = args.gamma_bndry_cons
        self.gamma_smthns_range = args.gamma_smthns
        self.gamma_smthns_cons_range = args.gamma_smthns_cons
        self.gamma_bndry_loc_range = args.gamma_bndry_loc
        self.gamma_depth_range = args.gamma_depth
        self.gamma_idx = -1
        self.process_group = process_group      # data-parallel: ranks hold disjoint slices of the global batch
        self._geo = _geometry_from_args(args)
        self.ctx = _lib.Context(_lib.make_config(max_batch=self.batch_size, **self._geo), self.device)
        self.H_patches, self.W_patches = self.ctx.Hp, self.ctx.Wp
        self.lambda_ridge = self.ctx.lambda_ridge
        self.global_image = self.global_bndry = self.terms = None

    _RANGES = ('gamma_color', 'gamma_color_cons', 'gamma_bndry_cons', 'gamma_smthns', 'gamma_smthns_cons', 'gamma_bndry_loc', 'gamma_depth')

    def calculate_gamma(self, gamma_range, rate, order=1):      # global_training.py:25-26
        return gamma_range[0] + rate ** order * (gamma_range[1] - gamma_range[0])

    def update_gamma(self, idx_update=True):                     # global_training.py:28-51
        if idx_update:
            self.gamma_idx += 1
        e0, e1, e2 = self.dynamic_epoch
        if self.gamma_idx < e0:
            rate, k = self.gamma_idx / (e0 - 1), 0
        elif self.gamma_idx < e1:
            rate, k = 1.0, 0
        elif self.gamma_idx < e2:
            rate, k = (self.gamma_idx - e1) / (e2 - e1 - 1), 1
        else:
            rate, k = 1.0, 1
        for name in self._RANGES:
            setattr(self, name, self.calculate_gamma(getattr(self, name + '_range')[k:k + 2], rate))

    def final_gamma(self):                                       # global_training.py:53-60
        for name in self._RANGES:
            setattr(self, name, getattr(self, name + '_range')[-1])

    def gammas(self):
        return [getattr(self, name) for name in self._RANGES]

    def _f32(self, t):
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    def forward(self, est, img_ny, img_gt, bndry_dist, deri, bndry_depth):
        """est [B,L,12] raw GlobalStage output; img_ny / img_gt [B,2,H,W,3]; bndry_dist / bndry_depth [B,H,W];
        deri [B,2,H-2,W-2,3] (global_training.py:147-157).  Returns the scalar loss (differentiable w.r.t. est)."""
        B, L = est.shape[0], self.ctx.L
        if est.dim() != 3 or est.shape[1:] != (L, 12):
            raise _lib.BlurryEdgesError(f'expects est [B,{L},12], got {tuple(est.shape)}')
        if B > self.ctx.max_batch:
            if torch.cuda.is_current_stream_capturing():      # a captured graph would keep pointers into the workspace freed here
                raise _lib.BlurryEdgesError(f'batch of {B} pairs exceeds the context ({self.ctx.max_batch}) inside a CUDA-graph capture: '
                                            'construct the criterion with batch_size >= the largest batch')
            self.ctx.close()
            self.ctx = _lib.Context(_lib.make_config(max_batch=B, **self._geo), self.device)
        raw = self._f32(est.detach())
        ny, gt = self._f32(img_ny), (img_ny if img_gt is img_ny else img_gt)
        gt = ny if gt is img_ny else self._f32(gt)
        want_grad = torch.is_grad_enabled() and est.requires_grad
        self.ctx.set_deterministic(want_deterministic(self.deterministic))
        args = (raw, ny, gt, self._f32(bndry_dist), self._f32(deri), self._f32(bndry_depth))
        npatch = B * L
        if self.process_group is None:
            gimg, gbnd, cnt = self.ctx.global_loss_stage1(*args)
            terms, loss, grad = self.ctx.global_loss_stage2(B, self.gammas(), npatch, cnt, want_grad)
        elif not self.overlap_collective:
            # plain recipe: the all-reduce sits between the two stages on the compute stream (its latency is exposed once per step)
            from .dist_utils import sync_loss_normalisers
            gimg, gbnd, cnt = self.ctx.global_loss_stage1(*args)
            cnt, assumed = sync_loss_normalisers(cnt, npatch, self.process_group)
            grad, gdep = self.ctx.global_loss_stage2_launch(B, self.gammas(), assumed, want_grad)
            terms, loss, grad = self.ctx.global_loss_stage2_finish(B, self.gammas(), assumed, cnt, grad, gdep)
        else:
            # Normalisers of the WHOLE batch: one 16-byte all-reduce of (mask count, patch count).  It is started as soon as the render
            # kernels are queued (the count is final when they finish): the collective's kernel becomes resident while the small
            # target kernels run - queued behind the loss kernel, which fills every SM, it would only get a slot in that kernel's
            # tail (measured at 8 GPUs: 3.25 ms per 32-pair step instead of 2.9) - and it runs while the loss kernel does; only the
            # final scalar combine and the depth term's share of the gradient wait for it.
            from .dist_utils import sync_loss_normalisers
            gimg, gbnd, cnt, (npatch, work) = self.ctx.global_loss_stage1(
                *args, between=lambda c: sync_loss_normalisers(c, npatch, self.process_group, async_op=True)[1:])
            grad, gdep = self.ctx.global_loss_stage2_launch(B, self.gammas(), npatch, want_grad)
            if work is not None:
                work.wait()
            terms, loss, grad = self.ctx.global_loss_stage2_finish(B, self.gammas(), npatch, cnt, grad, gdep)
        self.global_image, self.global_bndry, self.terms, self.mask_count = gimg, gbnd, terms, cnt[:1]
        self.last_loss_share = loss
        scale = 1.0
        if self.process_group is not None and self.grad_reduce == 'mean':
            import torch.distributed as dist
            scale = float(dist.get_world_size(self.process_group))
        if want_grad:
            return _FusedLossFn.apply(est, loss, grad.to(est.dtype), scale)
        return loss.reshape(()) * scale if scale != 1.0 else loss.reshape(())


class LocalLossFused(nn.Module):
    def __init__(self, args, device='cuda:0'):
        super().__init__()
        self.device = torch.device(device)
        self.R, self.w = int(args.R), float(args.w)
        self.batch_size = int(args.batch_size)
        self.max_beta_bndry_loc = args.beta_bndry_loc       # local_training.py:13-16
        self.max_beta_smthns = args.beta_smthns
        self.beta_idx = -1
        self.dynamic_epoch = args.dynamic_epoch
        self._geo = dict(R=self.R, stride=1, H=self.R, W=self.R, w=self.w, alpha_lambda=float(args.alpha_lambda),
                         cam=dict(args.cam_params), mag=float(args.mag))
        self.ctx = _lib.Context(_lib.make_config(max_batch=self.batch_size, **self._geo), self.device)
        self.lambda_ridge = self.ctx.lambda_ridge
        self.terms = None

    def update_beta(self, idx_update=True):                  # local_training.py:18-26
        if idx_update:
            self.beta_idx += 1
        rate = self.beta_idx / (self.dynamic_epoch - 1) if self.beta_idx < self.dynamic_epoch else 1.0
        self.beta_bndry_loc = rate * self.max_beta_bndry_loc
        self.beta_smthns = rate * self.max_beta_smthns

    def final_beta(self):                                    # local_training.py:28-30
        self.beta_bndry_loc = self.max_beta_bndry_loc
        self.beta_smthns = self.max_beta_smthns

    def _f32(self, t):
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    def forward(self, est, img_ny, gt_img, bndry_dist, deri):
        """est [B,10] raw LocalStage output; img_ny / gt_img [B,R,R,3]; bndry_dist [B,R,R]; deri [B,R-2,R-2,3]
        (local_training.py:47-52).  Like the reference (:33) the angles of `est` are wrapped IN PLACE."""
        B = est.shape[0]
        if est.dim() != 2 or est.shape[1] != 10:
            raise _lib.BlurryEdgesError(f'expects est [B,10], got {tuple(est.shape)}')
        if B > self.ctx.max_batch:
            self.ctx.close()
            self.ctx = _lib.Context(_lib.make_config(max_batch=B, **self._geo), self.device)
        want_grad = torch.is_grad_enabled() and est.requires_grad
        det = est.detach()
        # The reference mutates the network output (local_training.py:33: est[:, 4:8] = remainder(...)).  When est is a plain fp32
        # CUDA tensor that autograd allows to be written (not a leaf that requires grad) the kernel wraps the angles in place itself.
        in_place = det.is_cuda and det.dtype == torch.float32 and det.is_contiguous() and not (est.requires_grad and est.is_leaf)
        raw = det if in_place else self._f32(det)
        ny = self._f32(img_ny)
        gt = ny if gt_img is img_ny else self._f32(gt_img)
        terms, loss, grad = self.ctx.local_loss(raw, ny, gt, self._f32(bndry_dist), self._f32(deri), self.beta_bndry_loc,
                                                self.beta_smthns, want_grad, wrap_in_place=in_place)
        self.terms = terms
        if not in_place and not (est.requires_grad and est.is_leaf):
            with torch.no_grad():
                est.data[:, 4:8] = torch.remainder(est.data[:, 4:8], 2 * torch.pi)
        if want_grad:
            return _FusedLossFn.apply(est, loss, grad.to(est.dtype))
        return loss.reshape(())
