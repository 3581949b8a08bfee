"""Host-side mirror of the reference's helper classes (utils/postprocessing_loss.py, utils/depth_etas.py): same names,
constructor arguments, attributes and method signatures, every method backed by a kernel of this library (forward and
backward).  The reference's scripts subclass these; see tools/run_reference_script.py for running them unchanged."""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops


def _cam_cfg(args, H, W, stride, max_batch=1):
    return _lib.make_config(R=int(args.R), stride=int(stride), H=int(H), W=int(W), w=float(args.w), alpha_lambda=float(args.alpha_lambda),
                            cam=dict(args.cam_params), mag=float(args.mag), rho_prime=float(getattr(args, 'rho_prime', 10.39)),
                            max_batch=max_batch)


def _adjugate3(A):
    """adj(A) of [...,3,3] matrices by cofactors (device tensor ops, differentiable).  The reference builds it from A^2 and
    traces (utils/postprocessing_loss.py:127-128,148-149), which loses ~3e-3 in fp32; `inverse_3by3`
    here does not go through it (one kernel, fp64 cofactors), the method exists for subclasses that call it directly."""
    a, b, c = A[..., 0, 0], A[..., 0, 1], A[..., 0, 2]
    d, e, f = A[..., 1, 0], A[..., 1, 1], A[..., 1, 2]
    g, h, i = A[..., 2, 0], A[..., 2, 1], A[..., 2, 2]
    rows = [torch.stack([e * i - f * h, c * h - b * i, b * f - c * e], -1),
            torch.stack([f * g - d * i, a * i - c * g, c * d - a * f], -1),
            torch.stack([d * h - e * g, b * g - a * h, a * e - b * d], -1)]
    return torch.stack(rows, -2)


class DepthEtas:
    """utils/depth_etas.py:3-37"""

    def __init__(self, args, device):
        self.device = torch.device(device)
        self._be = _lib.Context(_cam_cfg(args, args.R, args.R, 1), self.device)
        self.s = args.cam_params['s']
        self.numerator = self._be.numerator
        self.denominator_constant = self._be.denominator_constant
        self.denominator_factor_root = self._be.denominator_factor_root
        self.denominator_factor = self._be.denominator_factor
        self.intercept = torch.tensor(self._be.intercept, dtype=torch.float32, device=self.device)
        self.theta_mid = torch.tensor(3 / 4 * torch.pi, device=self.device)
        self.theta_wng = torch.tensor(1 / 4 * torch.pi, device=self.device)

    def etas2depth(self, eta1, eta2):
        return ops.Etas2Depth.apply(eta1, eta2, self._be)

    def depth2sigma(self, depth, rho_prime):
        return ops.Elementwise.apply(depth, self._be, 2, rho_prime)


class PostProcessBase(nn.Module, ABC):
    """utils/postprocessing_loss.py:7-117"""

    def __init__(self, args, device):
        super().__init__()
        self.device = torch.device(device)
        self.R = args.R
        self.batch_size = args.batch_size
        self.w = args.w
        self.lambda_ridge = (args.alpha_lambda * self.R ** 2) ** 2
        H, W = self._image_size(args)
        self._be = _lib.Context(_cam_cfg(args, H, W, getattr(args, 'stride', 1)), self.device)
        yy, xx = torch.meshgrid([torch.linspace(-1.0, 1.0, self.R), torch.linspace(-1.0, 1.0, self.R)], indexing='ij')
        self.x, self.y = self.get_xy_mat(xx, yy)
        k = torch.tensor([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=torch.float32, device=self.device)
        self.sobel_x = k.unsqueeze(0).unsqueeze(0).repeat(3, 1, 1, 1)
        self.sobel_y = (-k.t()).unsqueeze(0).unsqueeze(0).repeat(3, 1, 1, 1)

    def _image_size(self, args):
        return args.R, args.R

    @abstractmethod
    def get_xy_mat(self, xx, yy):
        pass

    @abstractmethod
    def get_adjA(self, A, A2, trA, trA2):
        pass

    # Compatibility helpers of the reference's own params2dists (utils/postprocessing_loss.py:26-41): nothing in this package calls
    # them (params2dists is one kernel), they exist for subclasses that do.  Both are one coordinate of the pixel grid expressed in
    # the frame of an edge through (x, y) with direction `angle`.
    def _edge_frame(self, x, y, angle):
        """-> (normal coordinate, axial coordinate) of every pixel, differentiable torch ops"""
        s, c = torch.sin(angle), torch.cos(angle)
        dx, dy = self.x - x, self.y - y
        return torch.addcmul(c * dy, s, dx, value=-1.0), torch.addcmul(c * dx, s, dy)

    def dist4edge(self, x, y, angle):
        return self._edge_frame(x, y, angle)[0]

    def dist4axial(self, x, y, angle):
        return self._edge_frame(x, y, angle)[1]

    def itemize_params(self, params):
        return tuple(params[:, k, ...].unsqueeze(1).unsqueeze(1) for k in range(8))

    def params2dists(self, params):
        return ops.Params2Dists.apply(params, self._be, self.R)

    def params2etas(self, params):
        return ops.Elementwise.apply(params, self._be, 0, 0.0)

    def dists2indicators(self, dists, etas):
        return ops.Dists2Indicators.apply(dists, etas, self._be)

    def normalized_gaussian(self, x, delta=0.07):
        return ops.Elementwise.apply(x, self._be, 1, delta)

    def inverse_3by3(self, A):
        return ops.Inverse3.apply(A, self._be)

    def get_image_derivative(self, img):
        return ops.ImageDerivative.apply(img, self._be)


class PostProcessLocalBase(PostProcessBase):
    """utils/postprocessing_loss.py:119-128"""

    def __init__(self, args, device):
        super().__init__(args, device)
        self.ridge = self.lambda_ridge * torch.eye(3, device=self.device).unsqueeze(0)

    def get_xy_mat(self, xx, yy):
        return xx.view(1, self.R, self.R).to(self.device), yy.view(1, self.R, self.R).to(self.device)

    def get_adjA(self, A, A2, trA, trA2):
        return _adjugate3(A)


class PostProcessGlobalBase(PostProcessBase):
    """utils/postprocessing_loss.py:130-173"""

    def __init__(self, args, device):
        super().__init__(args, device)
        self.stride = args.stride
        self.H = args.img_size[0]
        self.W = args.img_size[1]
        self.ridge = self.lambda_ridge * torch.eye(3, device=self.device).unsqueeze(0).unsqueeze(0).unsqueeze(0)
        self.H_patches = int(np.floor((self.H - self.R) / self.stride) + 1)
        self.W_patches = int(np.floor((self.W - self.R) / self.stride) + 1)
        self.num_patches = self._be.cover_count()

    def _image_size(self, args):
        return args.img_size[0], args.img_size[1]

    def get_xy_mat(self, xx, yy):
        return xx.view(1, self.R, self.R, 1, 1).to(self.device), yy.view(1, self.R, self.R, 1, 1).to(self.device)

    def get_adjA(self, A, A2, trA, trA2):
        return _adjugate3(A)

    def _fold(self, patches, planes, mode=0):
        return ops.Fold.apply(patches, self._be, planes, self.H, self.W, mode)

    def local2global_color(self, patches, pair=True):
        B = self.batch_size
        if pair:
            return self._fold(patches, B * 6).view(B, 2, 3, self.H, self.W)
        return self._fold(patches, B * 3).view(B, 3, self.H, self.W)

    def local2global_bndry(self, bndry_patches):
        return self._fold(bndry_patches, self.batch_size).view(self.batch_size, 1, self.H, self.W)

    def local2global_depth(self, depth_map, depth_mask):
        B = self.batch_size
        dm = depth_map.to(torch.float32).contiguous()
        mk = depth_mask.to(torch.int32).contiguous()
        depth = torch.empty(B, self.H, self.W, device=dm.device, dtype=torch.float32)
        conf = torch.empty_like(depth)
        self._be.call('be_fold_depth', dm, mk, B, depth, conf)
        return depth, conf
