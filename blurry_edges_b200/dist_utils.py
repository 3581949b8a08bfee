"""The two collectives of the path, kept separate so that the CPU test-suite can run them under gloo."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def sync_loss_normalisers(counts: torch.Tensor, local_patches: int, group=None, async_op: bool = False):
    """Normalisers of the WHOLE batch of a data-parallel loss step, in ONE collective: counts = int64[2] on the device holding
    (depth-term mask count of this rank, patches of this rank), summed in place over `group` (global_training.py:127 divides by the
    mask count of the whole batch, the other six terms by the patch count of the whole batch).  A one-element tensor (mask count only)
    is accepted too.  Returns (counts, assumed_global_patches[, work]): assumed = local_patches * world is what the host can know
    without waiting for the collective - the kernels are launched with it and `global_loss_stage2_finish` rescales by
    assumed / counts[1] when the shards turn out to be uneven (a last batch without drop_last).  With async_op=True `work` is the
    pending all-reduce (None if there was nothing to do): the caller overlaps it with the loss kernel and calls work.wait() before
    it uses the counts."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return (counts, local_patches, None) if async_op else (counts, local_patches)
    world = dist.get_world_size(group)
    work = None
    if world > 1:
        work = dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return (counts, local_patches * world, work if async_op else None) if async_op else (counts, local_patches * world)


def row_bands(H: int, world: int):
    """Image rows owned by each rank: `world` contiguous bands of (almost) equal height."""
    per, rem = divmod(H, world)
    out, y = [], 0
    for r in range(world):
        n = per + (1 if r < rem else 0)
        out.append((y, y + n))
        y += n
    return out


def exchange_row_bands(acc: torch.Tensor, acc_rows, span, spans, bands, group=None):
    """Big-image block sharding, the one exchange step.  Every rank has folded its blocks into `acc` [rows, W, C], which holds image
    rows acc_rows = [r0, r1) - an interval that contains both span = [a, b), the rows its blocks wrote, and the band of rows the
    rank OWNS; spans / bands list every rank's span and band.  ONE all_to_all_single sends each other owner exactly the rows of its
    band that this rank wrote (consecutive owners -> contiguous slices of acc: no packing) and the received pieces are added IN PLACE
    to the rank's own rows of acc (fixed source order).  Returns the view acc[own band] - complete sums, ready to be normalised.
    A rank ships the few rows it shares with its neighbours instead of the whole image that a reduce onto one rank moved
    (67.5 MB at 1027 x 1027), and its own rows never move."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    W, C = acc.shape[-2], acc.shape[-1]
    cut = lambda sp, bd: (max(sp[0], bd[0]), max(min(sp[1], bd[1]), max(sp[0], bd[0])))       # rows of span sp inside band bd (may be empty)
    send = [cut(span, bands[o]) if o != rank else (span[0], span[0]) for o in range(world)]
    recv = [cut(spans[s], bands[rank]) if s != rank else (0, 0) for s in range(world)]
    in_split = [(y1 - y0) * W * C for y0, y1 in send]
    out_split = [(y1 - y0) * W * C for y0, y1 in recv]
    r0 = acc_rows[0]
    # rows sent to owners below this rank, then rows sent to owners above it: two contiguous runs of acc; all_to_all_single wants one
    # contiguous input in rank order, so the (at most two) runs are concatenated - a few shared rows, not the accumulator
    lowr = [(y0, y1) for o, (y0, y1) in enumerate(send) if o < rank and y1 > y0]
    high = [(y0, y1) for o, (y0, y1) in enumerate(send) if o > rank and y1 > y0]
    parts = []
    for runs in (lowr, high):
        if runs:
            parts.append(acc[min(y0 for y0, _ in runs) - r0:max(y1 for _, y1 in runs) - r0].reshape(-1))
    src = parts[0] if len(parts) == 1 else (torch.cat(parts) if parts else acc.new_empty(0))
    got = torch.empty(sum(out_split), dtype=acc.dtype, device=acc.device)
    dist.all_to_all_single(got, src.contiguous(), out_split, in_split, group=group)
    off = 0
    for (y0, y1), n in zip(recv, out_split):                                                  # fixed source order: rank 0, 1, ...
        if n:
            acc[y0 - r0:y1 - r0] += got[off:off + n].view(y1 - y0, W, C)
        off += n
    y0b, y1b = bands[rank]
    return acc[y0b - r0:y1b - r0]


def gather_row_bands(packed: torch.Tensor, bands, group=None, dst_group_rank: int = 0):
    """packed [P, band rows, W]: the finished planes of this rank's band -> [P, H, W] on one rank (None elsewhere).  ONE gather of
    one message per rank (bands padded to the tallest), then one concatenation that drops the padding."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    hmax = max(y1 - y0 for y0, y1 in bands)
    P, rows, W = packed.shape
    if rows != hmax:
        pad = packed.new_empty(P, hmax, W)
        pad[:, :rows] = packed
        packed = pad
    dst = dist.get_global_rank(group, dst_group_rank) if group is not None else dst_group_rank
    lst = [torch.empty_like(packed) for _ in range(world)] if rank == dst_group_rank else None
    dist.gather(packed.contiguous(), lst, dst=dst, group=group)
    if rank != dst_group_rank:
        return None
    return torch.cat([t[:, :y1 - y0] for t, (y0, y1) in zip(lst, bands)], 1)


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process to the CPUs next to GPU `device_index` (NVML CPU affinity, else the sysfs `local_cpulist` of the PCI device),
    intersected with the CPUs the process may use.  Call it BEFORE allocating pinned host buffers: they are then placed
    first-touch on the GPU's own NUMA node, so the H2D / D2H traffic of the host-buffer entry point (be_host_render_fold) does not
    cross the socket interconnect when one rank per GPU runs on a multi-socket box.  Never raises; returns what it did."""
    info = {'device': device_index, 'bound': False}
    try:
        allowed = os.sched_getaffinity(0)
    except (AttributeError, OSError):
        return info
    cpus = set()
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            uuid = uuid if uuid.startswith('GPU-') else 'GPU-' + uuid
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (max(allowed) // 64) + 1
        for w, bits in enumerate(pynvml.nvmlDeviceGetCpuAffinity(h, words)):
            cpus |= {64 * w + b for b in range(64) if (int(bits) >> b) & 1}
        info['source'] = 'nvml'
        try:
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
            with open(f'/sys/bus/pci/devices/{bus[-12:]}/numa_node') as f:
                info['numa_node'] = int(f.read())
        except Exception:
            pass
    except Exception:
        cpus = set()
    if not cpus:
        try:
            bus = torch.cuda.get_device_properties(device_index).pci_bus_id
            dom = torch.cuda.get_device_properties(device_index).pci_domain_id
            dev = torch.cuda.get_device_properties(device_index).pci_device_id
            with open(f'/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist') as f:
                for part in f.read().strip().split(','):
                    lo, _, hi = part.partition('-')
                    cpus |= set(range(int(lo), int(hi or lo) + 1))
            info['source'] = 'sysfs'
        except Exception:
            return info
    use = cpus & allowed
    if not use or use == allowed:
        info['cpus'] = len(allowed)
        return info                      # single node, or an affinity mask that does not overlap what we may use
    try:
        os.sched_setaffinity(0, use)
        info.update(bound=True, cpus=len(use))
    except OSError:
        pass
    return info
