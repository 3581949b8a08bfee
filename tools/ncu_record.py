#!/usr/bin/env python
"""Per-launch figures of one kernel from an `ncu --set full` capture -> profiles/ncu_records.json, the file bench.py reads
`roofline.traffic` and `sm_issue.warp_inst_per_patch_ncu` from (so that the bench line never carries pasted constants: the numbers
travel with the name of the capture and the command that produced it).

  python tools/ncu_record.py <rep> <kernel regex> <record key> <pairs in the captured launch> "<command that was profiled>"
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, regex, key, pairs, cmd = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--kernel-name', f'regex:{regex}'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, first = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, first))
    u = dict(zip(hdr, units))

    def val(name, want_unit=None):
        v = float(d[name].replace(',', ''))
        scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 'second': 1e3}
        return v * scale.get(u[name], 1.0)

    L = 4096
    rec = {
        'kernel': d['Kernel Name'], 'pairs': pairs, 'grid': d['Grid Size'], 'registers_per_thread': int(float(d['launch__registers_per_thread'])),
        'duration_ms_under_ncu': val('gpu__time_duration.sum'),
        'dram_bytes_per_launch': val('dram__bytes_read.sum') + val('dram__bytes_write.sum'),
        'dram_bytes_read': val('dram__bytes_read.sum'), 'dram_bytes_write': val('dram__bytes_write.sum'),
        'warp_inst_per_patch': float(d['smsp__inst_executed.sum'].replace(',', '')) / (pairs * L),
        'issue_active_pct': float(d['sm__issue_active.avg.pct_of_peak_sustained_elapsed']),
        'l1_smem_data_pipe_pct': float(d['l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed']),
        'fma_pipe_pct': float(d['sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']),
        'source': os.path.basename(rep) + ' (ncu --set full --clock-control none --import-source on)', 'command': cmd,
    }
    path = os.path.join(ROOT, 'profiles', 'ncu_records.json')
    allrec = json.load(open(path)) if os.path.exists(path) else {}
    allrec[key] = rec
    with open(path, 'w') as f:
        json.dump(allrec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == '__main__':
    main()
