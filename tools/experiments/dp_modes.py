import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch, torch.distributed as dist, bench
from blurry_edges_b200 import GlobalLossFused
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
raw, ny, gt, bd, deri, zg = [t.to(dev) for t in bench.train_inputs(B, rank * B, seed=200 + rank)]
raw.requires_grad_(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, kw in (('no collective', dict(process_group=None)), ('overlapped', dict(process_group=dist.group.WORLD)),
                 ('between stages', dict(process_group=dist.group.WORLD, overlap_collective=False)), ('overlapped', dict(process_group=dist.group.WORLD))):
    crit = GlobalLossFused(bench.loss_args(B), None, dev, **kw)
    crit.update_gamma()
    def step():
        raw.grad = None
        crit(raw, gt, gt, bd, deri, zg).backward()
    for _ in range(5):
        step()
    torch.cuda.synchronize(); dist.barrier()
    ev = []
    for i in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / 20], device=dev, dtype=torch.float64)
    mx = ms.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    mn = ms.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f'B={B} {name}: per-step ms max over ranks {float(mx):.3f}, min {float(mn):.3f}', flush=True)
    del crit
dist.destroy_process_group()
