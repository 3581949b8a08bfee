#!/usr/bin/env python
"""Benchmark of the Blurry-Edges render -> fold -> depth hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs B]

A step = one pass B (blurry_edges_test.py:81-100: render both images with shared ridge colours, sharpened and
refocused renders, boundary, depth mask/map, five folds) over one batch of synthetic 147x147 image pairs
(config[1] of BASELINE.json: 64 pairs = 262144 patches per GPU).  N > 1 (torchrun, one rank per GPU): every rank
processes its own batch of pairs (weak scaling, no data-path collective); time = max over ranks.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` the same metric through the host-buffer
C-ABI entry point with H2D/D2H copies inside the timed region, `roofline` the HBM roofline of the dominant kernel
(be_run3_kernel, timed with CUDA events on its own stream), `cpu_baseline` the oracle port timed on this box's cores.
`--impl reference` times the CPU port of the reference's eager PyTorch path (the reference itself is Python and is
not present on the GPU box; oracle/be_oracle.py is pinned to it by tests/golden)."""
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
METRIC = 'patches/sec render+fold+depth (BASELINE configs[1]: inference pass B; the fwd+bwd figures are under train_step*)'
UNIT = 'patches/s'
# algorithmic bytes per patch of pass B with pixels sourced from the image pair (SURVEY.md 8d / DESIGN.md 4):
# params 48 + pixels 2*3*147^2*4/4096 = 126.6 + outputs 15 planes*147^2*4/4096 = 316.5
ALGO_BYTES_PER_PATCH = 48.0 + 2 * 3 * S * S * 4 / L + 15 * S * S * 4 / L
# the same with pixels sourced from the unfolded [2,3,R,R,Hp,Wp] tensor the reference signature hands in (SURVEY.md 8d "bytes_api")
API_BYTES_PER_PATCH = 48.0 + 2 * 3 * R * R * 4 + 15 * S * S * 4 / L
# dram__bytes_read.sum + dram__bytes_write.sum of be_run3_kernel<INFER> for one 64-pair launch, from the committed
# `ncu --set full` capture profiles/r1y_run3_kernel_full.txt; None for other batch sizes
TRAFFIC_NCU_64 = 155.284992e6 + 38.297600e6
TRAFFIC_NCU = None
# warp-instructions per patch of be_run3_kernel<INFER> (smsp__inst_executed.sum / patches, profiles/r1y_run3_kernel_full.txt)
WARP_INST_PER_PATCH = 3868.0
SM_COUNT, SMSP_PER_SM = 148, 4


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 4 ms from a thread (the timed region of
    the default run lasts tens of milliseconds, too short for `nvidia-smi -lms`), nvidia-smi as the fallback."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.sm, self.mx, self.reasons, self.proc, self.stop = index, [], [], set(), None, False
        self.rows = []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                uuid = uuid if uuid.startswith('GPU-') else 'GPU-' + uuid
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, 'encode') else uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
            # the first NVML queries of a process take milliseconds (longer when several ranks ask at once): pay for them here,
            # not inside the timed region
            pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        bits = {'hw_slowdown': getattr(n, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                'hw_thermal_slowdown': getattr(n, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                'sw_thermal_slowdown': getattr(n, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                'sw_power_cap': getattr(n, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)}
        try:
            self.mx.append(int(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        except Exception:
            pass
        while True:                               # at least one sample, the last one taken after `stop` was requested
            last = self.stop
            try:
                self.sm.append(int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            if last:
                break
            time.sleep(0.004)

    def __enter__(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def __exit__(self, *a):
        self.stop = True
        if self.nvml is not None:
            self.thread.join(timeout=2)
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        if self.nvml is not None:
            sm = sorted(self.sm)
            return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(self.mx) if self.mx else None,
                    'reasons': sorted(self.reasons), 'samples': len(sm), 'source': 'nvml, 4 ms period, over the device-resident and the e2e timed regions'}
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm), 'source': 'nvidia-smi'}


def restore_params(raw):
    """The script glue between GlobalStage and the helper (blurry_edges_test.py:135-138): xy*3, angles wrapped to [0,2pi),
    eta coefficients + 0.5.  Input synthesis only; written here so that the product arm never imports oracle/."""
    import math
    return torch.cat([raw[..., :4] * 3, torch.remainder((raw[..., 4:8] + 1) * math.pi, 2 * math.pi), raw[..., 8:] + 0.5], dim=-1)


def make_inputs(B, seed):
    import synth
    est = restore_params(synth.raw_global(B, L, seed=seed)).contiguous()
    img = synth.image_pairs(B, S, S, seed=seed + 1).permute(0, 1, 4, 2, 3).contiguous()   # planar [B,2,3,H,W]
    return est, img


def cpu_reference_step(est, img, threads):
    """The reference's eager fp32 CPU path for pass B, restated (oracle/be_oracle.py, trace-formula inverse as in
    utils/postprocessing_loss.py:104-112); like the reference helper it handles one pair per call."""
    from oracle import be_oracle as O
    g, cam = O.Geometry(H=S, W=S), O.Camera()
    torch.set_num_threads(threads)
    with torch.no_grad():
        for b in range(est.shape[0]):
            O.inference(est[b:b + 1], img[b:b + 1], g, cam, 10.39, None, trace_form=True)


def time_cpu(pairs, reps, threads):
    est, img = make_inputs(pairs, seed=900)
    cpu_reference_step(est[:1], img[:1], threads)   # warm-up
    best = float('inf')
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_reference_step(est, img, threads)
        best = min(best, time.perf_counter() - t0)
    return pairs * L / best, best


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    pairs = args.ref_pairs
    est, img = make_inputs(pairs, seed=900)
    for _ in range(args.warmup):
        cpu_reference_step(est[:1], img[:1], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(est, img, threads)
    dt = time.perf_counter() - t0
    val = args.steps * pairs * L / dt
    out = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': f'pass B on {S}x{S} pairs, R={R}, stride={STRIDE}; each step = {pairs} pairs ({pairs * L} patches), '
                                  'a bounded sample of the 64-pair batch'},
           'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                            'sample': f'{pairs} pairs/step x {args.steps} steps, torch {torch.__version__} eager fp32 CPU, '
                                      'oracle/be_oracle.py restatement of the reference (one pair per call, as the reference helper)'},
           'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    try:   # the fwd+bwd side of the metric: GlobalLoss forward + autograd backward of the same CPU port, one pair (bounded sample)
        import synth
        from oracle import be_oracle as O
        g, cam = O.Geometry(H=S, W=S), O.Camera()
        raw = synth.raw_global(1, L, seed=300).requires_grad_(True)
        img_t = synth.image_pairs(1, S, S, seed=301)
        gt, bd, deri, zg = synth.loss_targets(1, S, S, seed=302)
        gam = [1.0, 0.2, 0.05, 0.005, 0.005, 1e-4, 1e-4]          # gamma_idx = 0 (global_training.py:28-51)
        t0 = time.perf_counter()
        O.global_loss(raw, img_t, gt, bd, deri, zg, gam, g, cam, trace_form=True).backward()
        dt2 = time.perf_counter() - t0
        out['train_step'] = {'metric': 'patches/sec loss fwd+bwd (GlobalLoss, CPU port, autograd)', 'value': L / dt2, 'unit': UNIT,
                             'sample': '1 pair, 1 repetition', 'seconds': dt2}
    except Exception as e:
        out['train_step'] = {'error': str(e)[:120]}
    _emit(out)


def _timed(fn, steps, warmup, dev, barrier, world):
    """ms per call of fn (CUDA events on the current stream, max over ranks)."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def extra_configs(args, rank, world, dev, barrier):
    """Secondary measurements of the other BASELINE.json configs (same JSON line, separate keys):
    configs[2] global-loss training step fwd+bwd (4 pairs per GPU = batch 32 on 8 GPUs), configs[3] one 1027x1027 pair with
    its 121 blocks sharded over the ranks, configs[4] densify 'w' with 32 pairs per GPU (256 on 8 GPUs)."""
    import argparse as ap
    import synth
    from blurry_edges_b200 import BigImageFused, Context, GlobalLossFused, PostProcessFused, _lib, make_config, shard_blocks
    import torch.distributed as dist
    cam = {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}
    base = dict(R=R, stride=STRIDE, w=1.0, alpha_lambda=5e-3, img_size=[S, S], mag=4.0, rho_prime=10.39, cam_params=cam)
    out = {}
    steps = max(3, args.steps // 2)
    # ---- configs[2]: training step of the loss (forward + analytic backward), data parallel -------------------------
    Bt = 4
    targs = ap.Namespace(batch_size=Bt, gamma_color=[1.0, 0.1, 0.1], gamma_color_cons=[0.2, 0.1, 0.05], gamma_bndry_cons=[0.05, 0.05, 0.02],
                         gamma_smthns=[0.005, 0.1, 0.002], gamma_smthns_cons=[0.005, 0.1, 0.002], gamma_bndry_loc=[0.0001, 0.05, 0.0001],
                         gamma_depth=[0.0001, 0.05, 0.5], dynamic_epoch=[30, 100, 200], **base)
    crit = GlobalLossFused(targs, None, dev, process_group=(dist.group.WORLD if world > 1 else None))
    crit.update_gamma()
    raw = synth.raw_global(Bt, L, seed=300 + rank).to(dev).requires_grad_(True)
    img = synth.image_pairs(Bt, S, S, seed=301 + rank).to(dev)
    gt, bd, deri, zg = [t.to(dev) for t in synth.loss_targets(Bt, S, S, seed=302 + rank)]

    def train_step():                                   # the training call of the reference passes the clean image twice
        raw.grad = None                                 # (global_training.py:210: criteria(est, img_gt, img_gt, ...))
        crit(raw, gt, gt, bd, deri, zg).backward()

    ms = _timed(train_step, steps, 3, dev, barrier, world)
    out['train_step'] = {'metric': 'patches/sec loss fwd+bwd (GlobalLoss, configs[2])', 'value': Bt * L * world / (ms / 1e3), 'unit': UNIT,
                         'ms_per_step': ms, 'pairs_per_gpu': Bt, 'collective': '8-byte mask-count all-reduce between the two loss stages',
                         'call': 'criteria(est, img_gt, img_gt, bndry_dist, deri, bndry_depth) + backward, as global_training.py:210-211'}
    # the same step with the whole batch of configs[2] (32 pairs) on every GPU: throughput of the kernels once the GPU is full
    Bt2 = 32
    targs.batch_size = Bt2
    crit2 = GlobalLossFused(targs, None, dev, process_group=(dist.group.WORLD if world > 1 else None))
    crit2.update_gamma()
    raw2 = synth.raw_global(Bt2, L, seed=330 + rank).to(dev).requires_grad_(True)
    img2 = synth.image_pairs(Bt2, S, S, seed=331 + rank).to(dev)
    gt2, bd2, deri2, zg2 = [t.to(dev) for t in synth.loss_targets(Bt2, S, S, seed=332 + rank)]

    def train_step2():
        raw2.grad = None
        crit2(raw2, gt2, gt2, bd2, deri2, zg2).backward()

    def val_step2():                                    # the validation call (global_training.py:166) has distinct noisy / clean images
        raw2.grad = None
        crit2(raw2, img2, gt2, bd2, deri2, zg2).backward()

    ms_val = _timed(val_step2, steps, 3, dev, barrier, world)
    ms = _timed(train_step2, steps, 3, dev, barrier, world)
    out['train_step_b32'] = {'metric': 'patches/sec loss fwd+bwd (GlobalLoss), 32 pairs per GPU', 'value': Bt2 * L * world / (ms / 1e3),
                             'unit': UNIT, 'ms_per_step': ms, 'pairs_per_gpu': Bt2, 'ms_per_step_distinct_noisy_and_clean_images': ms_val,
                             'kernels_ncu': {'be_loss2_kernel': {'ms': 2.18, 'warp_inst_per_patch': 9539, 'issue_active': 0.50, 'l1_smem_pipe': 0.68},
                                             'be_run3_kernel<TRAINFWD>': {'ms': 0.57, 'warp_inst_per_patch': 3198, 'issue_active': 0.65, 'l1_smem_pipe': 0.72},
                                             'source': 'profiles/r1y_train_kernels_full.txt'},
                             'algorithmic_bytes_per_patch': 810.0,
                             'hbm_frac_at_algorithmic_bytes': 810.0 * Bt2 * L / (ms / 1e3) / 1e9 / peaks()[0]}
    del crit2, raw2, img2, gt2, bd2, deri2, zg2
    # ---- the Smish activation of LocalStage, an HBM-bound elementwise kernel (8 B/element fwd, 12 B/element bwd) ----
    from blurry_edges_b200 import smish
    xs = torch.rand(8192, 64, R, R, device=dev) * 8 - 4            # one LocalStage activation of 4096 patches x 2 images: 925 MB
    gs = torch.rand_like(xs)
    ms_f = _timed(lambda: smish(xs), steps, 3, dev, barrier, world)
    xg = xs.clone().requires_grad_(True)

    def smish_fb():
        xg.grad = None
        smish(xg).backward(gs)

    ms_fb = _timed(smish_fb, steps, 3, dev, barrier, world)
    hbm, _ = peaks()
    nb = xs.numel() * 4
    out['smish'] = {'metric': 'Smish activation (models/local_stage.py:4-6), fused elementwise kernel', 'elements': xs.numel(),
                    'fwd_ms': ms_f, 'fwd_gbs': 2 * nb / (ms_f / 1e3) / 1e9, 'fwd_frac_of_hbm_peak': 2 * nb / (ms_f / 1e3) / 1e9 / hbm,
                    'fwd_bwd_ms': ms_fb, 'bwd_gbs': 3 * nb / ((ms_fb - ms_f) / 1e3) / 1e9,
                    'note': 'working set 1.85 GB >> L2; bwd figure = 3 arrays / (fwd+bwd - fwd) time, includes autograd glue'}
    del xs, gs, xg
    # ---- pass A (blurry_edges_test.py:125-128, colors_only=True): ridge colours of 2 x 64 single images, reported separately (8d) ----
    Ba = 64
    ctx_a = Context(make_config(H=S, W=S, max_batch=Ba), dev)
    p10 = restore_params(synth.raw_global(2 * Ba, L, seed=340 + rank))[..., :10].contiguous().to(dev)
    img_a = synth.image_pairs(Ba, S, S, seed=341 + rank).permute(0, 1, 4, 2, 3).reshape(2 * Ba, 3, S, S).contiguous().to(dev)
    lay_a = _lib.single_planar_layout(S, S)
    ms = _timed(lambda: ctx_a.colors(p10, img_a, lay_a, _lib.PARAMS_LOCAL10), steps, 3, dev, barrier, world)
    out['pass_a'] = {'metric': 'patches/sec pass A (ridge colours per single image), 128 images per GPU', 'value': 2 * Ba * L * world / (ms / 1e3),
                     'unit': 'single-image patches/s', 'ms_per_step': ms}
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
    out['local_train_step'] = {'metric': 'patches/sec LocalLoss fwd+bwd (local_training.py: 64 single patches per step)', 'value': Bl * world / (ms / 1e3),
                               'unit': 'single patches/s', 'ms_per_step': ms, 'patches_per_gpu': Bl,
                               'note': 'launch bound: 3 kernels + autograd glue for 64 patches'}
    del lcrit, le, lny, lgt, lbd, lderi
    # ---- configs[4]: densify 'w' ------------------------------------------------------------------------------------
    Bw = 32
    pargs = ap.Namespace(batch_size=Bw, densify='w', **base)
    helper = PostProcessFused(pargs, None, dev, as_numpy=False)
    est_w = restore_params(synth.raw_global(Bw, L, seed=310 + rank)).to(dev)
    img_w = synth.image_pairs(Bw, S, S, seed=311 + rank).to(dev)
    ms = _timed(lambda: helper(est_w, img_w, colors_only=False), steps, 3, dev, barrier, world)
    out['dense_w'] = {'metric': "patches/sec pass B, --densify 'w' (configs[4])", 'value': Bw * L * world / (ms / 1e3), 'unit': UNIT,
                      'ms_per_step': ms, 'pairs_per_gpu': Bw}
    # ---- configs[3]: one 1027x1027 pair, 121 blocks sharded over the ranks ------------------------------------------
    big = 1027
    bargs = ap.Namespace(batch_size=1, big_img_size=[big, big], n_margin_patch=10, densify=None, **base)
    bh = BigImageFused(bargs, None, dev, process_group=(dist.group.WORLD if world > 1 else None))
    lo, hi = shard_blocks(bh.nblk, rank, world)
    est_b = restore_params(synth.raw_global(hi - lo, L, seed=320 + rank)).to(dev)
    big_img = (torch.from_numpy(synth.photon_pairs(1, big, big, seed=321)).float() / 190.0).permute(0, 1, 4, 2, 3)[0].contiguous().to(dev)
    ms = _timed(lambda: bh(est_b, big_img), steps, 3, dev, barrier, world)
    npatch_big = ((big - R) // STRIDE + 1) ** 2
    out['big_1027'] = {'metric': 'patches/sec pass B + stitch + fold of one 1027x1027 pair (configs[3])', 'value': npatch_big / (ms / 1e3),
                       'unit': UNIT, 'ms_per_image': ms, 'blocks': bh.nblk, 'blocks_this_rank': hi - lo, 'scaling': 'strong',
                       'collective': 'sum-reduce of the [1027,1027,16] accumulator onto rank 0' if world > 1 else None}
    return out


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from blurry_edges_b200 import Context, _lib, make_config

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback); use --impl reference for the CPU port')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    B = args.pairs
    # one rank per GPU on a multi-socket box: keep the rank (and, first-touch, its pinned buffers) on the GPU's NUMA node
    from blurry_edges_b200.dist_utils import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local_rank) if (world > 1 and not args.no_numa) else {'bound': False}
    global TRAFFIC_NCU
    TRAFFIC_NCU = TRAFFIC_NCU_64 if B == 64 else None
    est_h, img_h = make_inputs(B, seed=100 + rank)
    est_h, img_h = est_h.pin_memory(), img_h.pin_memory()
    est, img = est_h.to(dev), img_h.to(dev)
    ctx = Context(make_config(H=S, W=S, max_batch=B), dev)
    layout = _lib.planar_layout(S, S)
    out = ctx.alloc_outputs(B, True)
    # the six maps PostProcess.forward returns (blurry_edges_test.py:100); the script thresholds the depth on the host (:144)
    host_out = ctx.host_render_fold(est_h, img_h, layout, want_thresholded=False)   # allocates pinned outputs + staging (untimed)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def step():
        ctx.render_fold(est, img, layout, out=out)

    ctx.set_timing(True)
    for _ in range(args.warmup):
        step()
    barrier()
    n0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern = []
    with ClockSampler(local_rank) as clk:
        t_wall = time.perf_counter()
        for a, b in ev:
            flush.zero_()                                          # L2 flush between timed iterations (untimed)
            a.record()
            step()
            b.record()
            kern.append(ctx.last_timing())                         # waits for this step's last kernel
        barrier()
        t_wall = time.perf_counter() - t_wall
        launches = _lib.launch_count() - n0
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        ctx.set_timing(False)

        # end to end through the host-buffer C-ABI call: pinned host inputs -> H2D -> kernels -> D2H of the maps
        # (still inside the clock sampler: both timed regions are covered)
        for _ in range(max(1, args.warmup // 2)):
            ctx.host_render_fold(est_h, img_h, layout, out=host_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.host_render_fold(est_h, img_h, layout, out=host_out)   # synchronises internally
        barrier()
        e2e_s = time.perf_counter() - t0

    extra = {} if args.no_extra else extra_configs(args, rank, world, dev, barrier)

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()
    if rank != 0:
        return
    patches = B * L * world
    value = patches * args.steps / (dev_ms / 1e3)
    e2e = patches * args.steps / (e2e_ms / 1e3)
    run_ms = sum(k[2] for k in kern) / len(kern)
    shares = [sum(k[i] for k in kern) / len(kern) for i in range(4)]
    peak, peak_src = peaks()
    achieved = ALGO_BYTES_PER_PATCH * B * L / (run_ms / 1e3) / 1e9
    res = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
           'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': f'inference pass B (render+fold+depth), {B} synthetic {S}x{S} pairs per GPU = {B * L} patches/step/GPU, '
                                  f'R={R}, stride={STRIDE}, densify=None (BASELINE.json configs[1])',
                      'l2': 'flushed between timed iterations (256 MiB write, untimed); working set 217 MB > 126 MB L2',
                      'sharding': 'one batch per rank, no data-path collective'},
           'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(B * (L * 12 + 6 * S * S) * 4),
                   'd2h_bytes_per_step': int(B * 15 * S * S * 4), 'ms_per_step': e2e_ms / args.steps,
                   'api': 'be_host_render_fold (pinned host buffers, synchronous): est + image pair in, the six maps of '
                          'PostProcess.forward (15 planes, blurry_edges_test.py:100) out',
                   'numa_binding_rank0': numa},
           'gpu_launches': int(launches),
           'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                        'traffic': TRAFFIC_NCU, 'kernel': 'be_run3_kernel<INFER>', 'kernel_ms': run_ms, 'peak_source': peak_src,
                        'algorithmic_bytes_per_patch': ALGO_BYTES_PER_PATCH,
                        'frac_if_pixels_came_unfolded': API_BYTES_PER_PATCH * B * L / (run_ms / 1e3) / 1e9 / peak,
                        'note': 'the fused path is FP32/SFU-issue bound, not HBM bound (DESIGN.md section 4); '
                                'see profiles/ for pipe utilisation'},
           'sm_issue': {'note': 'binding resource of the fused kernel: warp-instruction issue slots (1 per SMSP per clock)',
                        'warp_inst_per_patch_ncu': WARP_INST_PER_PATCH,
                        'achieved_ginst_per_s': WARP_INST_PER_PATCH * B * L / (run_ms / 1e3) / 1e9,
                        'peak_ginst_per_s': SM_COUNT * SMSP_PER_SM * (clk.summary()['sm_mhz'] or 1965) / 1e3,
                        'frac': WARP_INST_PER_PATCH * B * L / (run_ms / 1e3) / (SM_COUNT * SMSP_PER_SM * (clk.summary()['sm_mhz'] or 1965) * 1e6)},
           'kernel_ms': {'memset': shares[0], 'be_setup_kernel': shares[1], 'be_run3_kernel': shares[2], 'be_normalise_kernel': shares[3]},
           'clocks': clk.summary(), 'wall_s_timed_region': t_wall}
    res.update(extra)
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, sec = time_cpu(args.ref_pairs, 2, cores)
        res['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                               'sample': f'{args.ref_pairs} of the {B} pairs (best of 2, {sec:.2f} s), torch eager fp32 CPU port of the '
                                         'reference (oracle/be_oracle.py), one pair per call'}
        try:
            from oracle import be_oracle as O, hostmath
            g, cam = O.Geometry(H=S, W=S), O.Camera()
            e2, i2 = make_inputs(8, seed=901)
            hostmath.render_fold(e2[:1], i2[:1], g, cam)
            t0 = time.perf_counter()
            hostmath.render_fold(e2, i2, g, cam)
            res['cpu_baseline']['c_port_openmp'] = {'value': 8 * L / (time.perf_counter() - t0), 'unit': UNIT,
                                                    'note': 'oracle/be_hostmath.cpp: the kernels\' arithmetic on host cores, OpenMP over pairs'}
        except Exception as e:  # the C port is optional evidence
            res['cpu_baseline']['c_port_openmp'] = {'error': str(e)[:100]}
        try:   # the reference's eager op chain on this same B200 (the oracle port on cuda:0, one pair per call)
            from oracle import be_oracle as O
            g, cam = O.Geometry(H=S, W=S), O.Camera()
            e3, i3 = make_inputs(2, seed=902)
            e3, i3 = e3.to(dev), i3.to(dev)
            with torch.no_grad(), torch.device(dev):     # the oracle's constants (grids, ridge) are made on the default device
                O.inference(e3[:1], i3[:1], g, cam, 10.39, None, trace_form=True)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                for b in range(2):
                    O.inference(e3[b:b + 1], i3[b:b + 1], g, cam, 10.39, None, trace_form=True)
                torch.cuda.synchronize(dev)
            res['cpu_baseline']['eager_port_on_this_gpu'] = {'value': 2 * L / (time.perf_counter() - t0), 'unit': UNIT,
                                                             'note': 'the same torch restatement of the reference run eagerly on cuda:0 (fp32, '
                                                                     'one pair per call): the launch/HBM-bound path the fused kernels replace'}
        except Exception as e:
            res['cpu_baseline']['eager_port_on_this_gpu'] = {'error': str(e)[:100]}
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
    ap.add_argument('--pairs', type=int, default=64, help='image pairs per GPU per step (BASELINE configs[1]: 64)')
    ap.add_argument('--ref-pairs', type=int, default=8, help='pairs per step of the CPU reference sample')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary configs (train step, densify w, big image)')
    ap.add_argument('--no-numa', action='store_true', help='multi-rank runs: do not bind each rank to the CPUs of its GPU\'s NUMA node')
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
