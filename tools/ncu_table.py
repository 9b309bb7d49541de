#!/usr/bin/env python
"""Summarise the CSV logs of tools/ncu_hbm_kernels.sh: per kernel (and grid size class) the median duration, DRAM bytes read +
written, L2 bytes and the throughput percentages ncu reports — as a markdown table for profiles/.
    python tools/ncu_table.py gpurun_out/r2_ncu > profiles/r02_hbm_kernels_ncu.md"""
import csv
import statistics
import sys
from collections import defaultdict
from pathlib import Path

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'msecond': 1.0, 'ms': 1.0, 'nsecond': 1e-6,
        'second': 1e3, 's': 1e3, '%': 1.0, 'register/thread': 1.0}


def load(path):
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    return rows


def main():
    root = Path(sys.argv[1])
    for path in sorted(root.glob('*.csv')):
        per = defaultdict(lambda: defaultdict(list))
        for r in load(path):
            try:
                val = float(r['Metric Value'].replace(',', '')) * UNIT.get(r['Metric Unit'], 1.0)
            except (ValueError, KeyError):
                continue
            name = r['Kernel Name'].split('(')[0].replace('srf::', '').replace('void ', '')
            key = (name, r.get('Grid Size', ''), r.get('Block Size', ''))
            per[key][r['Metric Name']].append((int(r['ID']), val))
        print(f'\n### {path.stem}\n')
        print('| kernel | grid | launches | time ms | DRAM read MB | DRAM write MB | DRAM GB/s | L2 MB | DRAM % | L2 % | SM % | regs | warps active % |')
        print('|---|---|---|---|---|---|---|---|---|---|---|---|---|')
        for key in sorted(per, key=lambda k: -statistics.median(v for _, v in per[k]['gpu__time_duration.sum'])):
            m = per[key]

            def med(metric):
                vals = [v for _, v in m.get(metric, [])]
                return statistics.median(vals) if vals else float('nan')
            t = med('gpu__time_duration.sum')
            rd, wr = med('dram__bytes_read.sum'), med('dram__bytes_write.sum')
            print(f"| {key[0]} | {key[1]} | {len(m['gpu__time_duration.sum'])} | {t:.4f} | {rd / 1e6:.2f} | {wr / 1e6:.2f} | {(rd + wr) / t / 1e6:.0f} | "
                  f"{med('lts__t_bytes.sum') / 1e6:.1f} | {med('dram__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                  f"{med('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {med('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                  f"{med('launch__registers_per_thread'):.0f} | {med('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} |")


if __name__ == '__main__':
    main()
