#!/usr/bin/env python
"""Per-source-line and per-opcode instruction counts of one kernel from an .ncu-rep (ncu --set full --import-source on).

  python tools/ncu_lines.py gpurun_out/prof_X.ncu-rep <patches in the captured launch> [top N]
"""
import csv
import subprocess
import sys
from collections import Counter


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    fname, hdr, lines, ops, osmp = None, None, [], Counter(), Counter()
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]
        elif r[0] == 'Line No':
            hdr = r
            iE, iS, iA = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Address')
        elif hdr and len(r) == len(hdr):
            if r[0].strip().isdigit():
                lines.append((fname, int(r[0]), r[1], int(r[iE] or 0), int(r[iS] or 0)))
            elif r[iA].startswith('0x'):
                t = r[iA + 1].split()
                op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0].rstrip(';')
                ops[op] += int(r[iE] or 0)
                osmp[op] += int(r[iS] or 0)
    tot, ts = sum(l[3] for l in lines), max(1, sum(l[4] for l in lines))
    print(f'warp-instructions per unit: {tot / units:.1f}')
    for op, n in ops.most_common(25):
        print(f'  {op:10s} {n / units:8.1f} {n / tot * 100:5.1f}%   samples {osmp[op] / ts * 100:4.1f}%')
    for l in sorted(lines, key=lambda l: -l[3])[:top]:
        print(f'{l[0]:12s}:{l[1]:4d} {l[3] / units:7.1f} {l[3] / tot * 100:5.1f}% smp {l[4] / ts * 100:4.1f}% | {l[2].strip()[:100]}')


if __name__ == '__main__':
    main()
