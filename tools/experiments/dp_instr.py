import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch, torch.distributed as dist, bench
from blurry_edges_b200 import GlobalLossFused
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
B = 32
host = [t.contiguous().pin_memory() for t in bench.train_inputs(B, rank * B, seed=200 + rank)]
raw, ny, gt, bd, deri, zg = [t.to(dev) for t in host]
raw.requires_grad_(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
crit = GlobalLossFused(bench.loss_args(B), None, dev, process_group=dist.group.WORLD)
crit.update_gamma()
def step():
    raw.grad = None
    crit(raw, gt, gt, bd, deri, zg).backward()
class Null:
    def __enter__(self): return self
    def __exit__(self, *a): pass
for name, timing, sampler in (('plain', False, False), ('lib timing events', True, False), ('sampler on rank 0', False, True), ('both (bench.py)', True, True), ('plain', False, False)):
    crit.ctx.set_timing(timing)
    for _ in range(3):
        step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev = []
    with (bench.ClockSampler(rank, active=(rank == 0)) if sampler else Null()):
        for i in range(20):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record()
            ev.append((a, b))
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / 20], device=dev, dtype=torch.float64)
    allms = [torch.zeros_like(ms) for _ in range(world)]
    dist.all_gather(allms, ms)
    if rank == 0:
        print(f'{name}: per-rank ms/step ' + ' '.join(f'{float(t):.3f}' for t in allms), flush=True)
crit.ctx.set_timing(False)
dist.destroy_process_group()
