import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch, bench
from blurry_edges_b200 import GlobalLossFused
B = 32
dev = torch.device('cuda:0')
host = [t.contiguous().pin_memory() for t in bench.train_inputs(B, 0, seed=200)]
raw_h, ny_h, gt_h, bd_h, deri_h, zg_h = host
crit = GlobalLossFused(bench.loss_args(B), None, dev)
crit.update_gamma()
g = crit.gammas()
out = crit.ctx.host_global_loss(raw_h, gt_h, gt_h, bd_h, deri_h, zg_h, g)
for _ in range(3):
    crit.ctx.host_global_loss(raw_h, gt_h, gt_h, bd_h, deri_h, zg_h, g, out=out)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    crit.ctx.host_global_loss(raw_h, gt_h, gt_h, bd_h, deri_h, zg_h, g, out=out)
torch.cuda.synchronize()
print(f"WGT={os.environ.get('BE_HOST_TRAIN_WGT')}: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per call")
