#!/usr/bin/env python
"""Headline metrics + the SASS instructions with the most warp-stall samples of the first kernel in an .ncu-rep.

  python tools/ncu_stalls.py gpurun_out/prof_X.ncu-rep <units in the captured launch> [top N]
"""
import csv
import subprocess
import sys


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = dict(zip(rows[0], rows[2]))
    for k in ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
              'launch__waves_per_multiprocessor', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
              'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
              'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
              'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
              'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active']:
        print(f'{k} = {d.get(k)}')
    print(f"warp-instructions per unit = {float(d['smsp__inst_executed.sum']) / units:.1f}")
    for k in rows[0]:
        if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio') and float(d[k] or 0) > 0.1:
            print(f'  stall {k[34:-23]:24s} {float(d[k]):.2f}')
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    iS, iN, iE = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(int(r[iN]) for r in data)
    for i in sorted(range(len(data)), key=lambda i: -int(data[i][iN]))[:top]:
        r = data[i]
        st = sorted(((hdr[c][6:], int(r[c] or 0)) for c in cols), key=lambda kv: -kv[1])[:3]
        print(f'{i:5d} {int(r[iN]) / tot * 100:5.2f}% ex/unit={int(r[iE]) / units:6.2f} {r[iS].strip()[:64]:64s} {[s for s in st if s[1]]}')


if __name__ == '__main__':
    main()
