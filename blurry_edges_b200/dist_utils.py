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


def reduce_accumulator(acc: torch.Tensor, group=None, dst_group_rank: int = 0):
    """Sum the ranks' partial fold accumulators ([1,H,W,16]) onto one rank (big-image block sharding)."""
    if dist.get_world_size(group) > 1:
        dst = dist.get_global_rank(group, dst_group_rank) if group is not None else dst_group_rank
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return acc


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
