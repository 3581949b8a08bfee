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
raw, ny, gt, bd, deri, zg = [t.to(dev) for t in bench.train_inputs(B, rank * B, seed=200 + rank)]
raw.requires_grad_(True)
crit = GlobalLossFused(bench.loss_args(B), None, dev, process_group=dist.group.WORLD)
crit.update_gamma()
def step():
    raw.grad = None
    crit(raw, gt, gt, bd, deri, zg).backward()
for _ in range(5):
    step()
torch.cuda.synchronize(); dist.barrier()
for delay_ms, sync_each in ((0.0, True), (1.0, True), (0.0, False), (1.0, False)):
    ts = []
    torch.cuda.synchronize(); dist.barrier()
    t_all = time.perf_counter()
    for i in range(20):
        if rank == 1 and delay_ms:
            time.sleep(delay_ms / 1e3)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        if sync_each:
            b.synchronize()
        ts.append((a, b))
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t_all) / 20 * 1e3
    ms = sum(a.elapsed_time(b) for a, b in ts) / 20
    print(f'rank {rank}: rank-1 delay {delay_ms} ms, host sync each step {sync_each}: {ms:.3f} ms per step (events), {t_all:.3f} ms wall per step', flush=True)
    dist.barrier()
dist.destroy_process_group()
