"""The inference driver body of the reference (blurry_edges_test.depth_estimator, blurry_edges_test.py:114-149) with every
step that is not a network on this library's kernels: patch gather -> LocalStage -> pass A (+ angle wrap) -> pm assembly ->
GlobalStage -> pass B (restore inside the kernel) -> confidence threshold -> depth metrics.  The two networks stay stock
PyTorch modules supplied by the caller.  With `cuda_graph=True` the whole body (networks included) is captured once per input
shape into a CUDA graph and replayed: at the reference's batch size of 1 the body is launch bound."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .fused import _geometry_from_args


class DepthEstimatorFused(nn.Module):
    def __init__(self, args, local_module, global_module, device='cuda:0', max_batch=None, cuda_graph=False):
        super().__init__()
        self.cuda_graph = bool(cuda_graph)
        self._graphs = {}
        self.device = torch.device(device)
        self.local_module, self.global_module = local_module, global_module
        self.densify = getattr(args, 'densify', None)
        self.crop = int(getattr(args, 'crop', 10))
        self.R, self.H, self.W = int(args.R), int(args.img_size[0]), int(args.img_size[1])
        self._geo = _geometry_from_args(args)
        self.ctx = _lib.Context(_lib.make_config(max_batch=int(max_batch or args.batch_size), **self._geo), self.device)
        self.L = self.ctx.L

    @torch.no_grad()
    def forward(self, img_ny, gt_depth=None):
        """img_ny [B,2,H,W,3] (dataset-native, already divided by alpha) -> dict(image, sharp, refoc, bndry, depth, conf,
        depth_map [B,H,W] thresholded as blurry_edges_test.py:144, metrics [B,5] if gt_depth [B,H,W] is given).
        With cuda_graph=True the returned tensors are the graph's static outputs: they are overwritten by the next call."""
        if not self.cuda_graph:
            return self._run(img_ny, gt_depth)
        key = (tuple(img_ny.shape), gt_depth is not None)
        entry = self._graphs.get(key)
        if entry is None:
            # Captured graphs bake in the context's workspace pointers (patch table, accumulator): size the context for this batch
            # BEFORE warm-up / capture, and throw away every graph captured against a context that is about to be replaced.
            self._ensure_batch(img_ny.shape[0])
            s_img = torch.empty(tuple(img_ny.shape), device=self.device, dtype=torch.float32)
            s_gt = None if gt_depth is None else torch.empty(tuple(gt_depth.shape), device=self.device, dtype=torch.float32)
            s_img.copy_(img_ny)
            if s_gt is not None:
                s_gt.copy_(gt_depth)
            cur, side = torch.cuda.current_stream(self.device), torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):          # warm-up outside the capture: workspace allocation, cuDNN plans, kernel attributes
                for _ in range(2):
                    self._run(s_img, s_gt)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._run(s_img, s_gt)
            entry = self._graphs[key] = (graph, s_img, s_gt, out)
        graph, s_img, s_gt, out = entry
        s_img.copy_(img_ny, non_blocking=True)
        if s_gt is not None:
            s_gt.copy_(gt_depth, non_blocking=True)
        graph.replay()
        return out

    def _ensure_batch(self, B):
        """Grow the context to B pairs.  A replaced context frees its workspace, so graphs captured against it are dropped (replaying
        them would read and write freed device memory); growing inside a stream capture is refused."""
        if B <= self.ctx.max_batch:
            return
        if torch.cuda.is_current_stream_capturing():
            raise _lib.BlurryEdgesError(f'batch of {B} pairs exceeds the context ({self.ctx.max_batch}) inside a CUDA-graph capture: '
                                        'construct DepthEstimatorFused with max_batch >= the largest batch')
        self._graphs.clear()
        torch.cuda.synchronize(self.device)              # replays of the dropped graphs may still be in flight
        self.ctx.close()
        self.ctx = _lib.Context(_lib.make_config(max_batch=B, **self._geo), self.device)

    def _run(self, img_ny, gt_depth=None):
        B, H, W, R, L = img_ny.shape[0], self.H, self.W, self.R, self.L
        self._ensure_batch(B)
        planar = img_ny.to(device=self.device, dtype=torch.float32).permute(0, 1, 4, 2, 3).contiguous()      # [B,2,3,H,W]
        vec = torch.empty(2 * B * L, 3, R, R, device=self.device, dtype=torch.float32)
        self.ctx.call('be_patch_gather', planar, 2 * B, vec)                                                  # :119-121
        params = self.local_module(vec).to(torch.float32).reshape(2 * B, L, 10).contiguous()                  # :122-123
        colors = self.ctx.colors(params, planar.view(2 * B, 3, H, W), _lib.single_planar_layout(H, W), _lib.PARAMS_LOCALRAW10)  # :125-128
        pm = torch.empty(B, L, 38, device=self.device, dtype=torch.float32)
        self.ctx.call('be_assemble_pm', params, colors, B, pm)                                                # :129-132
        raw = self.global_module(pm).to(torch.float32).contiguous()                                           # :134
        out = self.ctx.render_fold(raw, planar, _lib.planar_layout(H, W), densify_w=(self.densify == 'w'),
                                   param_mode=_lib.PARAMS_RAW12)                                              # :135-144
        res = dict(zip(('image', 'sharp', 'refoc', 'bndry', 'depth', 'conf', 'depth_map'), out))
        if gt_depth is not None:
            gt = gt_depth.to(device=self.device, dtype=torch.float32).contiguous()
            sums = torch.empty(B, 6, device=self.device, dtype=torch.float64)
            self.ctx.call('be_eval_depth', out[6], gt, B, H, W, self.crop, sums)                              # :148-149, utils/metrics.py
            n = sums[:, 0:1]
            res['metrics'] = torch.cat([sums[:, 1:4] / n, torch.sqrt(sums[:, 4:5] / n) * 100, sums[:, 5:6] / n * 100], 1)
        return res
