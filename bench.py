#!/usr/bin/env python
"""Benchmark of the Blurry-Edges render -> fold -> depth hot path, forward + backward (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs B]

A step = one global-stage training step of the hot path (global_training.py:208-211 without the transformer and the optimiser):
GlobalLoss.forward (split_restore_params, render both images with shared ridge colours, boundary map, analytic depth, the two
folds, the seven loss terms) + backward to the network output `est`, on one batch of basic-shape image pairs of the reference's
own generator (BASELINE.json configs[2]; 32 pairs = 131072 patches per GPU).  The call is the reference's training call
criteria(est, img_gt, img_gt, bndry_dist, deri, bndry_depth): the clean image is passed twice (global_training.py:210).
N > 1 (torchrun, one rank per GPU): every rank holds its own 32 pairs of the global batch of 32 N (weak scaling); the only
exchange is the 16-byte all-reduce of (depth-mask count, patch count) between the two kernel stages, inside the timed region.

Prints ONE JSON line (rank 0).  `value`: device-resident throughput through GlobalLossFused + backward; `e2e`: the same step
through the host-buffer C-ABI entry point (be_host_global_loss: pinned host inputs -> H2D -> kernels -> D2H of loss + grad est),
copies inside the timed region; `roofline`: HBM roofline of the dominant kernel (be_loss2_kernel, timed live with CUDA events on its
stream); `cpu_baseline`: the CPU port of the reference's eager path (oracle/be_oracle.py + autograd) on this box's cores.
`--impl reference` times that CPU port alone (the reference is Python source under /root/reference, absent on the GPU box; the
port is pinned to it by tests/golden and was timed side by side with it)."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

S, R, STRIDE = 147, 21, 2
HP = (S - R) // STRIDE + 1
L = HP * HP
METRIC = 'patches/sec render+fold+depth fwd+bwd (GlobalLoss training step on est, BASELINE configs[2])'
UNIT = 'patches/s'
# algorithmic bytes per patch of the training step, API-native layouts: est 48 + grad 48 + img_ny 126.6 + img_gt 126.6 +
# bndry_dist 21.1 + deri 123.2 + bndry_depth 21.1 + folded global image / boundary written once and read once 2 x 147.7
TRAIN_BYTES_PER_PATCH = 48 + 48 + 2 * (2 * 3 * S * S * 4 / L) + 2 * (S * S * 4 / L) + 2 * 3 * (S - 2) * (S - 2) * 4 / L + 2 * (7 * S * S * 4 / L)
# inference pass B (configs[1]), pixels sourced from the image pair: params 48 + pixels 126.6 + 15 output planes 316.5
INFER_BYTES_PER_PATCH = 48.0 + 2 * 3 * S * S * 4 / L + 15 * S * S * 4 / L
SM_COUNT, SMSP_PER_SM = 148, 4
CAM = {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}
GAMMA_RANGES = dict(gamma_color=[1.0, 0.1, 0.1], gamma_color_cons=[0.2, 0.1, 0.05], gamma_bndry_cons=[0.05, 0.05, 0.02],
                    gamma_smthns=[0.005, 0.1, 0.002], gamma_smthns_cons=[0.005, 0.1, 0.002], gamma_bndry_loc=[0.0001, 0.05, 0.0001],
                    gamma_depth=[0.0001, 0.05, 0.5], dynamic_epoch=[30, 100, 200])       # utils/args.py:53-61
GAMMAS_IDX0 = [1.0, 0.2, 0.05, 0.005, 0.005, 1e-4, 1e-4]                                  # update_gamma() once (global_training.py:28-51)


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_record(kernel):
    """Per-launch figures of one `ncu --set full` capture, as committed under profiles/ by tools/ncu_record.py (never measured in
    this run: the source file and its command travel with the numbers; None if the file is missing)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_records.json')) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions by a separate `nvidia-smi -lms 20` PROCESS watching the reporting
    rank's GPU (started before the warm-up, so that its first sample is in before the timed region begins; rows are time-stamped by
    a reader thread that sleeps in read()).  Nothing in the benchmark process itself talks to NVML while steps are being timed:
    at 8 GPUs a Python thread polling NVML every 4 ms made the polling rank's host fall behind, and through the per-step
    collective every other rank showed 3.55 ms per step instead of 2.85; five inline polls in 20 steps cost 1.4 ms per step
    (tools/experiments/dp_instr.py, profiles/r2_experiments.txt)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0, active=True):
        self.index, self.active, self.rows, self.proc, self.t0, self.t1 = index, active, [], None, None, None
        if not active:
            return
        try:
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            sel = uuid if uuid.startswith('GPU-') else 'GPU-' + uuid           # CUDA_VISIBLE_DEVICES may renumber: select by UUID
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20', '-i', sel],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t = time.perf_counter()
            while not self.rows and time.perf_counter() - t < 5.0:              # nvidia-smi needs ~1 s to produce its first row
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(',')]))

    def __enter__(self):
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        self.t1 = time.perf_counter()

    def close(self):
        if self.proc:
            time.sleep(0.03)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            self.proc = None

    def summary(self):
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) + 0.02]
        used, where = (inside, 'inside') if inside else ([r for _, r in self.rows[-2:]], 'nearest to')
        sm = sorted(int(r[0]) for r in used if r and r[0].isdigit())
        mx = [int(r[1]) for r in used if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith('active') for r in used)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons, 'samples': len(sm),
                'source': f'nvidia-smi -lms 20 in its own process on the reporting rank\'s GPU; samples {where} the device-resident + e2e timed regions'}


def restore_params(raw):
    """The script glue between GlobalStage and the helper (blurry_edges_test.py:135-138): xy*3, angles wrapped to [0,2pi),
    eta coefficients + 0.5.  Input synthesis only; written here so that the product arm never imports oracle/."""
    import math
    return torch.cat([raw[..., :4] * 3, torch.remainder((raw[..., 4:8] + 1) * math.pi, 2 * math.pi), raw[..., 8:] + 0.5], dim=-1)


def train_inputs(B, first, seed):
    """One batch of the training step: raw network output est_raw = 0.1 * N(0,1)-like and basic-shape scenes
    of the reference's generator (tests/golden/shapes147.npz, 8 scenes used round-robin; throughput does not depend on the values)."""
    import synth
    raw = synth.raw_global(B, L, seed=seed)
    ny, gt, bd, deri, zg = synth.shapes_batch(B, first=first)
    return raw, ny, gt, bd, deri, zg


def loss_args(B):
    return argparse.Namespace(R=R, stride=STRIDE, w=1.0, alpha_lambda=5e-3, img_size=[S, S], mag=4.0, rho_prime=10.39, cam_params=CAM,
                              batch_size=B, **GAMMA_RANGES)


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's eager fp32 path for the training step, restated (oracle/be_oracle.py + autograd; trace-formula inverse
# as in utils/postprocessing_loss.py:104-112)
# ---------------------------------------------------------------------------------------------------------------------
def bench_config(B):
    """`config` of the JSON line.  BOTH arms print this same dict (the reference arm measures a bounded sample of this workload and says
    which in `cpu_baseline.sample`), so that the driver's same-config check of the two arms holds."""
    return {'workload': f'GlobalLoss training step fwd+bwd (BASELINE.json configs[2]): {B} basic-shape {S}x{S} pairs per GPU = {B * L} '
                        f'patches/step/GPU, R={R}, stride={STRIDE}, gamma_idx 0, est_raw = 0.1 N(0,1); scenes from the reference generator '
                        '(tests/golden/shapes147.npz, 8 scenes round-robin)',
            'call': 'criteria(est, img_gt, img_gt, bndry_dist, deri, bndry_depth) + backward (global_training.py:210-211)',
            'l2': 'GPU arm: flushed between timed iterations (256 MiB write, untimed)',
            'sharding': 'GPU arm: one 32-pair slice of the global batch per rank, 16-byte all-reduce of (mask count, patch count) inside the step; '
                        'reference arm: rank 0 only, each step a bounded sample of the workload (cpu_baseline.sample)'}


def cpu_train_step(inp, threads):
    from oracle import be_oracle as O
    raw, ny, gt, bd, deri, zg = inp
    g, cam = O.Geometry(H=S, W=S), O.Camera()
    torch.set_num_threads(threads)
    r = raw.clone().requires_grad_(True)
    O.global_loss(r, gt, gt, bd, deri, zg, GAMMAS_IDX0, g, cam, trace_form=True).backward()     # the training call: clean image twice
    return r.grad


def time_cpu(pairs, warmup, reps, threads):
    inp = train_inputs(pairs, 0, seed=900)
    for _ in range(warmup):
        cpu_train_step(inp, threads)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_train_step(inp, threads)
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    pairs = args.ref_pairs
    ts = time_cpu(pairs, args.warmup, args.steps, threads)
    dt = sum(ts)
    val = args.steps * pairs * L / dt
    out = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': bench_config(args.pairs),
           'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                            'sample': f'{pairs} pair(s)/step x {args.steps} steps after {args.warmup} warm-up steps, torch {torch.__version__} eager fp32 '
                                      'CPU + autograd, oracle/be_oracle.py restatement of the reference (timed within 3 % of the unmodified '
                                      'reference in the build container)',
                            'best_step_s': min(ts)},
           'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    _emit(out)


def _timed(fn, steps, warmup, dev, barrier, world, flush=None):
    """ms per call of fn (CUDA events on the current stream, max over ranks); with `flush`, L2 is flushed (untimed) between calls."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    barrier()
    if flush is None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b) / steps
    else:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            flush.zero_()
            a.record()
            fn()
            b.record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _graph_of(fn, dev):
    """fn captured into a CUDA graph after two warm-up calls on a side stream (collectives included: NCCL kernels are capturable)."""
    cur, side = torch.cuda.current_stream(dev), torch.cuda.Stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        fn()
        fn()
    cur.wait_stream(side)
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def parity_checks(rank, world, dev):
    """N > 1 only: the sharded computations against the single-rank ones on the same global inputs (every rank can build all of
    them; the inputs are functions of the global sample index)."""
    import torch.distributed as dist
    import synth
    from blurry_edges_b200 import BigImageFused, GlobalLossFused, shard_blocks
    out = {}
    # ---- training: ranks hold uneven slices (rank r holds 1 + (r % 2) pairs) of one global batch ----
    sizes = [1 + (r % 2) for r in range(world)]
    lo = sum(sizes[:rank])
    Bg = sum(sizes)
    raw, ny, gt, bd, deri, zg = [t.to(dev) for t in train_inputs(Bg, 0, seed=700)]
    full = GlobalLossFused(loss_args(Bg), None, dev)
    full.update_gamma()
    rf = raw.clone().requires_grad_(True)
    lf = full(rf, gt, gt, bd, deri, zg)
    lf.backward()
    shard = GlobalLossFused(loss_args(sizes[rank]), None, dev, process_group=dist.group.WORLD, grad_reduce='sum')
    shard.update_gamma()
    sl = slice(lo, lo + sizes[rank])
    rs = raw[sl].clone().requires_grad_(True)
    ls = shard(rs, gt[sl], gt[sl], bd[sl], deri[sl], zg[sl])
    ls.backward()
    tot = ls.detach().clone()
    dist.all_reduce(tot)
    e = torch.stack([(tot - lf.detach()).abs() / lf.detach().abs(), (rs.grad - rf.grad[sl]).abs().max() / rf.grad.abs().max()]).double()
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    out['train'] = {'loss_rel_err_sum_of_shares_vs_single_rank': float(e[0]), 'grad_max_err_rel_to_max': float(e[1]),
                    'pairs_per_rank': sizes, 'ok': bool(e[0] < 2e-6 and e[1] < 2e-6)}
    del full, shard
    # ---- big image: 323x323 (3x3 blocks) sharded over the ranks against all blocks on one rank ----
    big = 323
    bargs = argparse.Namespace(R=R, stride=STRIDE, w=1.0, alpha_lambda=5e-3, img_size=[S, S], mag=4.0, rho_prime=10.39, cam_params=CAM,
                               batch_size=1, big_img_size=[big, big], n_margin_patch=10, densify=None)
    img = (torch.from_numpy(synth.photon_pairs(1, big, big, seed=63)).float() / 190.0).permute(0, 1, 4, 2, 3)[0].contiguous().to(dev)
    est = torch.stack([restore_params(synth.raw_global(1, L, seed=170 + k))[0] for k in range(9)]).to(dev)
    sh = BigImageFused(bargs, None, dev, process_group=dist.group.WORLD)
    b0, b1 = shard_blocks(sh.nblk, rank, world)
    maps = sh(est[b0:b1], img)
    worst = torch.zeros(1, dtype=torch.float64, device=dev)
    if rank == 0:
        single = BigImageFused(bargs, None, dev)(est, img)
        worst[0] = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(maps, single))
    dist.broadcast(worst, 0)
    out['big'] = {'maps_max_err_rel_to_max_sharded_vs_single_rank': float(worst), 'size': big, 'ok': bool(worst < 2e-6)}
    return out


def extra_configs(args, rank, world, dev, barrier, flush):
    """Secondary measurements of the other BASELINE.json configs (same JSON line, separate keys)."""
    import argparse as ap
    import synth
    from blurry_edges_b200 import BigImageFused, Context, GlobalLossFused, PostProcessFused, _lib, make_config, shard_blocks
    import torch.distributed as dist
    base = dict(R=R, stride=STRIDE, w=1.0, alpha_lambda=5e-3, img_size=[S, S], mag=4.0, rho_prime=10.39, cam_params=CAM)
    out = {}
    steps = max(3, args.steps // 2)
    pg = dist.group.WORLD if world > 1 else None
    # ---- configs[2] exactly: a global batch of 32 pairs split over the ranks (strong scaling: 32 / N pairs per GPU) ----
    if 32 % world == 0:
        Bt = 32 // world
        crit = GlobalLossFused(loss_args(Bt), None, dev, process_group=pg)
        crit.update_gamma()
        raw, ny, gt, bd, deri, zg = [t.to(dev) for t in train_inputs(Bt, rank * Bt, seed=300 + rank)]
        raw.requires_grad_(True)

        def strong_step():
            raw.grad = None
            crit(raw, gt, gt, bd, deri, zg).backward()

        ms = _timed(strong_step, steps, 3, dev, barrier, world, flush)
        out['train_step_batch32_total'] = {'metric': 'patches/sec fwd+bwd, global batch 32 split over the ranks (configs[2] as written)',
                                           'value': 32 * L / (ms / 1e3), 'unit': UNIT, 'ms_per_step': ms, 'pairs_per_gpu': Bt, 'scaling': 'strong'}
        # the validation call (global_training.py:166) has distinct noisy / clean images
        def val_step():
            raw.grad = None
            crit(raw, ny, gt, bd, deri, zg).backward()

        ms = _timed(val_step, steps, 3, dev, barrier, world, flush)
        out['train_step_batch32_total']['ms_per_step_distinct_noisy_and_clean_images'] = ms
        if args.graphs:   # the same step (all-reduce included) replayed as one CUDA graph: what a captured training loop pays
            try:
                gr = _graph_of(strong_step, dev)
                ms_g = _timed(gr.replay, steps, 3, dev, barrier, world, flush)
                out['train_step_batch32_total'].update(ms_per_step_cuda_graph=ms_g, value_cuda_graph=32 * L / (ms_g / 1e3))
                del gr
            except Exception as e:
                out['train_step_batch32_total']['cuda_graph_error'] = str(e)[:120]
        del crit, raw, ny, gt, bd, deri, zg
    # ---- deterministic fold mode (torch.use_deterministic_algorithms, global_training.py:177): cost of the fixed-order fold ----
    try:
        Bd = 8
        critd = GlobalLossFused(loss_args(Bd), None, dev, deterministic=True)
        critd.update_gamma()
        raw, ny, gt, bd, deri, zg = [t.to(dev) for t in train_inputs(Bd, 0, seed=360 + rank)]
        raw.requires_grad_(True)

        def det_step():
            raw.grad = None
            critd(raw, gt, gt, bd, deri, zg).backward()

        ms_d = _timed(det_step, steps, 3, dev, barrier, world, flush)
        critd.deterministic = False
        ms_n = _timed(det_step, steps, 3, dev, barrier, world, flush)
        out['deterministic_fold'] = {'pairs_per_gpu': Bd, 'ms_per_step': ms_d, 'ms_per_step_atomic_fold': ms_n}
        del critd, raw, ny, gt, bd, deri, zg
    except TypeError:
        pass
    # ---- configs[1]: inference pass B on 64 pairs per GPU (blurry_edges_test.py:81-100), device-resident and through host buffers ----
    Bi = 64
    est_h = restore_params(synth.raw_global(Bi, L, seed=100 + rank)).contiguous().pin_memory()
    img_h = synth.image_pairs(Bi, S, S, seed=101 + rank).permute(0, 1, 4, 2, 3).contiguous().pin_memory()
    est, img = est_h.to(dev), img_h.to(dev)
    ctx = Context(make_config(H=S, W=S, max_batch=Bi), dev)
    layout = _lib.planar_layout(S, S)
    o = ctx.alloc_outputs(Bi, True)
    ctx.set_timing(True)
    ms = _timed(lambda: ctx.render_fold(est, img, layout, out=o), steps, 3, dev, barrier, world, flush)
    kt = ctx.last_timing()
    ctx.set_timing(False)
    host_out = ctx.host_render_fold(est_h, img_h, layout, want_thresholded=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.host_render_fold(est_h, img_h, layout, out=host_out)
    barrier()
    e2e_ms = (time.perf_counter() - t0) / steps * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    hbm, _ = peaks()
    out['inference_b64'] = {'metric': 'patches/sec inference pass B (configs[1]), 64 pairs per GPU', 'value': Bi * L * world / (ms / 1e3), 'unit': UNIT,
                            'ms_per_step': ms, 'kernel_ms_be_run3': kt[2], 'e2e_value': Bi * L * world / (float(t.item()) / 1e3),
                            'hbm_frac_of_kernel_at_algorithmic_bytes': INFER_BYTES_PER_PATCH * Bi * L / (kt[2] / 1e3) / 1e9 / hbm}
    del ctx, est, img, est_h, img_h, o, host_out
    # ---- the Smish activation of LocalStage, an HBM-bound elementwise kernel (8 B/element fwd) ----
    from blurry_edges_b200 import smish
    xs = torch.rand(8192, 64, R, R, device=dev) * 8 - 4            # one LocalStage activation of 4096 patches x 2 images: 925 MB
    ms_f = _timed(lambda: smish(xs), steps, 3, dev, barrier, world)
    nb = xs.numel() * 4
    out['smish'] = {'elements': xs.numel(), 'fwd_ms': ms_f, 'fwd_gbs': 2 * nb / (ms_f / 1e3) / 1e9, 'fwd_frac_of_hbm_peak': 2 * nb / (ms_f / 1e3) / 1e9 / hbm}
    del xs
    # ---- pass A (blurry_edges_test.py:125-128, colors_only=True): ridge colours of 2 x 64 single images ----
    Ba = 64
    ctx_a = Context(make_config(H=S, W=S, max_batch=Ba), dev)
    p10 = restore_params(synth.raw_global(2 * Ba, L, seed=340 + rank))[..., :10].contiguous().to(dev)
    img_a = synth.image_pairs(Ba, S, S, seed=341 + rank).permute(0, 1, 4, 2, 3).reshape(2 * Ba, 3, S, S).contiguous().to(dev)
    lay_a = _lib.single_planar_layout(S, S)
    ms = _timed(lambda: ctx_a.colors(p10, img_a, lay_a, _lib.PARAMS_LOCAL10), steps, 3, dev, barrier, world)
    out['pass_a'] = {'value': 2 * Ba * L * world / (ms / 1e3), 'unit': 'single-image patches/s', 'ms_per_step': ms}
    del ctx_a, p10, img_a
    # ---- local-stage training step (local_training.py:99-108: LocalLoss forward + backward, 64 patches per step in the reference) ----
    from blurry_edges_b200 import LocalLossFused
    Bl = 64
    largs = ap.Namespace(batch_size=Bl, beta_bndry_loc=0.001, beta_smthns=0.0005, dynamic_epoch=200, **base)   # utils/args.py:34-36
    lcrit = LocalLossFused(largs, dev)
    lcrit.final_beta()
    le, lny, lgt, lbd, lderi = [t.to(dev) for t in synth.local_batch(Bl, R, seed=350 + rank)]
    le.requires_grad_(True)

    def local_step():
        le.grad = None
        lcrit(le, lny, lgt, lbd, lderi).backward()

    ms = _timed(local_step, max(steps, 10), 3, dev, barrier, world)
    out['local_train_step'] = {'value': Bl * world / (ms / 1e3), 'unit': 'single patches/s', 'ms_per_step': ms, 'patches_per_gpu': Bl,
                               'note': 'one kernel launch per step (setup, loss, backward, reduction fused); eager = Python + autograd glue bound'}
    try:   # the same step captured in a CUDA graph (what a captured training loop pays for it)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            local_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        le.grad = None
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            lcrit(le, lny, lgt, lbd, lderi).backward()
        out['local_train_step']['ms_per_step_cuda_graph'] = _timed(gr.replay, max(steps, 10), 3, dev, barrier, world)
        del gr
    except Exception as e:
        out['local_train_step']['cuda_graph_error'] = str(e)[:100]
    del lcrit, le, lny, lgt, lbd, lderi
    # ---- configs[4]: densify 'w' ----
    Bw = 32
    pargs = ap.Namespace(batch_size=Bw, densify='w', **base)
    helper = PostProcessFused(pargs, None, dev, as_numpy=False)
    est_w = restore_params(synth.raw_global(Bw, L, seed=310 + rank)).to(dev)
    img_w = synth.image_pairs(Bw, S, S, seed=311 + rank).to(dev)
    ms = _timed(lambda: helper(est_w, img_w, colors_only=False), steps, 3, dev, barrier, world)
    out['dense_w'] = {'value': Bw * L * world / (ms / 1e3), 'unit': UNIT, 'ms_per_step': ms, 'pairs_per_gpu': Bw}
    del helper, est_w, img_w
    # ---- configs[3]: one 1027x1027 pair, 121 blocks sharded over the ranks ----
    big = 1027
    bargs = ap.Namespace(batch_size=1, big_img_size=[big, big], n_margin_patch=10, densify=None, **base)
    bh = BigImageFused(bargs, None, dev, process_group=pg)
    lo, hi = shard_blocks(bh.nblk, rank, world)
    est_b = restore_params(synth.raw_global(hi - lo, L, seed=320 + rank)).to(dev)
    big_img = (torch.from_numpy(synth.photon_pairs(1, big, big, seed=321)).float() / 190.0).permute(0, 1, 4, 2, 3)[0].contiguous().to(dev)
    ms = _timed(lambda: bh(est_b, big_img), steps, 3, dev, barrier, world)
    npatch_big = ((big - R) // STRIDE + 1) ** 2
    out['big_1027'] = {'value': npatch_big / (ms / 1e3), 'unit': UNIT, 'ms_per_image': ms, 'blocks': bh.nblk, 'blocks_this_rank': hi - lo,
                       'scaling': 'strong'}
    if world > 1:
        ms_s = _timed(lambda: bh(est_b, big_img, gather=False), steps, 3, dev, barrier, world)
        out['big_1027'].update(collective='all_to_all of row bands (each rank owns 1/N of the image rows) + gather of the finished bands onto rank 0',
                               ms_per_image_maps_left_sharded_by_rows=ms_s, value_maps_left_sharded_by_rows=npatch_big / (ms_s / 1e3))
    if args.graphs:       # at 8 ranks a rank renders 16 blocks in 0.18 ms: the ~15 host-side calls of the eager path take longer than that
        try:
            for key, kw in (('cuda_graph', {}), ('cuda_graph_sharded_by_rows', {'gather': False})):
                if kw and world == 1:
                    continue
                gr = _graph_of(lambda: bh(est_b, big_img, **kw), dev)
                ms_g = _timed(gr.replay, steps, 3, dev, barrier, world)
                out['big_1027'][f'ms_per_image_{key}'] = ms_g
                out['big_1027'][f'value_{key}'] = npatch_big / (ms_g / 1e3)
                del gr
        except Exception as e:
            out['big_1027']['cuda_graph_error'] = str(e)[:120]
    return out


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from blurry_edges_b200 import GlobalLossFused, _lib

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback); use --impl reference for the CPU port')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    B = args.pairs
    from blurry_edges_b200.dist_utils import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local_rank) if (world > 1 and args.numa) else {'bound': False}
    pg = dist.group.WORLD if world > 1 else None
    host = [t.contiguous().pin_memory() for t in train_inputs(B, rank * B, seed=200 + rank)]
    raw_h, ny_h, gt_h, bd_h, deri_h, zg_h = host
    raw, gt, bd, deri, zg = [t.to(dev) for t in (raw_h, gt_h, bd_h, deri_h, zg_h)]
    raw.requires_grad_(True)
    crit = GlobalLossFused(loss_args(B), None, dev, process_group=pg)
    crit.update_gamma()                                             # gamma_idx 0, as the first epoch of global_training.py:205
    gammas = crit.gammas()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def step():                                                     # global_training.py:210-211 on est
        raw.grad = None
        crit(raw, gt, gt, bd, deri, zg).backward()

    clk = ClockSampler(local_rank, active=(rank == 0))            # its nvidia-smi process starts sampling now, before the warm-up
    crit.ctx.set_timing(True)
    for _ in range(args.warmup):
        step()
    barrier()
    n0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern = []
    with clk:
        t_wall = time.perf_counter()
        for i, (a, b) in enumerate(ev):
            flush.zero_()                                          # L2 flush between timed iterations (untimed)
            a.record()
            step()
            b.record()                                             # no host synchronisation between steps (a training loop has none)
        barrier()
        t_wall = time.perf_counter() - t_wall
        kern = [crit.ctx.last_train_timing(k) for k in range(min(args.steps, 64))]       # per-kernel CUDA events of the timed steps
        launches = _lib.launch_count() - n0
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        crit.ctx.set_timing(False)
        loss_val = float(crit.last_loss_share.item())

        # end to end through the host-buffer C-ABI call: pinned host inputs -> H2D -> kernels -> D2H of loss + grad est
        # (still inside the clock sampler: both timed regions are covered)
        hout = crit.ctx.host_global_loss(raw_h, gt_h, gt_h, bd_h, deri_h, zg_h, gammas, process_group=pg)
        for _ in range(max(1, args.warmup // 2)):
            crit.ctx.host_global_loss(raw_h, gt_h, gt_h, bd_h, deri_h, zg_h, gammas, out=hout, process_group=pg)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            crit.ctx.host_global_loss(raw_h, gt_h, gt_h, bd_h, deri_h, zg_h, gammas, out=hout, process_group=pg)   # synchronises internally
        barrier()
        e2e_s = time.perf_counter() - t0
    clk.close()
    # the two paths must agree on the result
    e2e_loss = float(hout[1].item())
    e2e_grad_err = float((hout[2].to(dev) - raw.grad / (world if pg is not None else 1)).abs().max() / (raw.grad.abs().max() / (world if pg is not None else 1)))

    parity = parity_checks(rank, world, dev) if world > 1 else None
    extra = {} if args.no_extra else extra_configs(args, rank, world, dev, barrier, flush)

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()
    if rank != 0:
        return
    patches = B * L * world
    value = patches * args.steps / (dev_ms / 1e3)
    e2e = patches * args.steps / (e2e_ms / 1e3)
    names = ['memset', 'be_setup_kernel', 'be_run3_kernel<TRAINFWD>', '(normalise: fused into be_train_targets_kernel)', 'be_train_targets_kernel', 'be_loss2_kernel',
             'reduce+fixup']
    shares = [sum(k[i] for k in kern) / len(kern) for i in range(7)]
    loss2_ms = shares[5]
    peak, peak_src = peaks()
    achieved = TRAIN_BYTES_PER_PATCH * B * L / (loss2_ms / 1e3) / 1e9
    rec = ncu_record('be_loss2_kernel') or {}
    clocks = clk.summary()
    h2d = int(sum(t.numel() for t in (raw_h, gt_h, bd_h, deri_h, zg_h)) * 4)
    res = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
           'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'f32', 'data': 'synthetic',
           'config': bench_config(B),
           'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': int(B * L * 12 * 4 + 32),
                   'ms_per_step': e2e_ms / args.steps,
                   'api': ('be_host_global_loss (two-phase schedule' if world == 1 else 'be_host_global_loss_begin/_end (deferred normaliser, 16-byte all-reduce between the halves')
                          + '; pinned host buffers, synchronous): est, clean image pair (passed twice, copied once), '
                          'bndry_dist, deri, bndry_depth in; terms, loss and grad est out',
                   'loss': e2e_loss, 'grad_max_err_vs_device_resident_path': e2e_grad_err, 'numa_binding_rank0': numa},
           'gpu_launches': int(launches),
           'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                        'traffic': rec.get('dram_bytes_per_launch') if B == rec.get('pairs') else None, 'traffic_source': rec.get('source'),
                        'kernel': 'be_loss2_kernel', 'kernel_ms': loss2_ms, 'peak_source': peak_src,
                        'algorithmic_bytes_per_patch': TRAIN_BYTES_PER_PATCH,
                        'note': 'the fused step is bound by the SM (issue slots, L1/shared-memory data pipe), not by HBM: DESIGN.md section 3.2'},
           'kernel_ms': dict(zip(names, shares)),
           'kernel_share_of_step': {'be_loss2_kernel': loss2_ms / (dev_ms / args.steps), 'be_run3_kernel<TRAINFWD>': shares[2] / (dev_ms / args.steps)},
           'loss_share_rank0': loss_val,
           'clocks': clocks, 'wall_s_timed_region': t_wall}
    if rec:
        wi = rec.get('warp_inst_per_patch')
        if wi:
            peak_inst = SM_COUNT * SMSP_PER_SM * (clocks['sm_mhz'] or 1965) * 1e6
            res['sm_issue'] = {'warp_inst_per_patch_ncu': wi, 'source': rec.get('source'),
                               'frac_of_issue_slots_at_measured_kernel_time': wi * B * L / (loss2_ms / 1e3) / peak_inst}
    if parity is not None:
        res['parity_check'] = parity
    res.update(extra)
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        ts = time_cpu(args.ref_pairs, 1, 3, cores)
        res['cpu_baseline'] = {'value': args.ref_pairs * L / (sum(ts) / len(ts)), 'unit': UNIT, 'cores': cores, 'kind': 'port',
                               'sample': f'{args.ref_pairs} of the {B} pairs per step, 1 warm-up + 3 timed steps ({sum(ts):.1f} s), torch eager fp32 CPU + '
                                         'autograd port of the reference (oracle/be_oracle.py)', 'best_step_s': min(ts)}
        try:   # the reference's eager op chain on this same B200 (the oracle port on cuda:0 with autograd)
            from oracle import be_oracle as O
            g, cam = O.Geometry(H=S, W=S), O.Camera()
            inp = [t.to(dev) for t in train_inputs(1, 0, seed=902)]
            with torch.device(dev):
                def eager():
                    r = inp[0].clone().requires_grad_(True)
                    O.global_loss(r, inp[2], inp[2], inp[3], inp[4], inp[5], GAMMAS_IDX0, g, cam, trace_form=True).backward()
                eager()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                for _ in range(3):
                    eager()
                torch.cuda.synchronize(dev)
            res['cpu_baseline']['eager_port_on_this_gpu'] = {'value': 3 * L / (time.perf_counter() - t0), 'unit': UNIT,
                                                             'note': 'the same torch restatement + autograd run eagerly on cuda:0 (fp32, 1 pair per step)'}
        except Exception as e:
            res['cpu_baseline']['eager_port_on_this_gpu'] = {'error': str(e)[:100]}
    # The driver keeps the parsed contract keys plus the last ~1500 characters of the line: put what is NOT a contract key but carries the
    # evidence (strong-scaling case, per-kernel times, the N>1 parity check) at the end.
    for k in ('train_step_batch32_total', 'big_1027', 'kernel_ms', 'kernel_share_of_step', 'parity_check'):
        if k in res:
            res[k] = res.pop(k)
    _emit(res)


_REAL_STDOUT = None


def _emit(obj):
    """The ONE JSON line, on the real stdout (everything else, e.g. NCCL's version banner, was redirected to stderr)."""
    line = (json.dumps(obj) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)      # libraries (NCCL) print to fd 1: keep stdout for the JSON line only
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--pairs', type=int, default=32, help='image pairs per GPU per step (BASELINE configs[2]: batch 32)')
    ap.add_argument('--ref-pairs', type=int, default=1, help='pairs per step of the CPU reference sample')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary configs (inference, densify w, big image, ...)')
    ap.add_argument('--graphs', action='store_true', help='also time CUDA-graph replays of the communicating secondary configs')
    ap.add_argument('--numa', action='store_true', help='multi-rank runs: bind each rank to the CPUs NVML lists for its GPU (off by default: on the 8-GPU box it '
                    'squeezed several ranks\' host threads onto few cores and cost 0.3 ms per step)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
