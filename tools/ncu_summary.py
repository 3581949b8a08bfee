#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_X.csv  profiles/X_launches.txt
  python tools/ncu_summary.py full     gpurun_out/prof_X.ncu-rep  profiles/X_full.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second']


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    d = defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ki]].append(float(r[vi].replace(',', '')))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    with open(dst, 'w') as f:
        f.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n# source: {src}\n')
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f'{sum(v) / tot * 100:6.2f}%  n={len(v):4d}  avg={sum(v) / len(v) / 1e3:10.1f} us  total={sum(v) / 1e6:9.3f} ms  {k[:110]}\n')
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full --clock-control none --import-source on; source: {src}\n')
        for vals in rows[2:]:
            rec = dict(zip(hdr, vals))
            f.write(f"\n## {rec.get('Kernel Name', '?')}  grid={rec.get('Grid Size')} block={rec.get('Block Size')}\n")
            for h, u, v in zip(hdr, units, vals):
                if h in KEYS or h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
                    f.write(f'{h} [{u}] = {v}\n')
    print(open(dst).read())


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
