#!/usr/bin/env python
"""Summarise .ncu-rep captures (read here, no GPU needed) into one CSV for profiles/.

    python tools/ncu_summary.py out.csv rep1.ncu-rep [rep2.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum']


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    w = None
    with open(out, 'w', newline='') as f:
        for rep in reps:
            txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(txt)))
            hdr, units = rows[0], rows[1]
            idx = {k: hdr.index(k) for k in KEYS if k in hdr}
            if w is None:
                w = csv.writer(f)
                w.writerow(['capture', 'kernel', 'grid'] + ['%s [%s]' % (k, units[idx[k]]) for k in KEYS if k in idx])
            ki, gi = hdr.index('Kernel Name'), hdr.index('Grid Size')
            for r in rows[2:]:
                w.writerow([rep.split('/')[-1], r[ki][:70], r[gi]] + [r[idx[k]] for k in KEYS if k in idx])


if __name__ == '__main__':
    main()
