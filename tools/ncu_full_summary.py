#!/usr/bin/env python
"""Key metrics of the `ncu --set full` reports of tools/ncu_full_top_kernels.sh as one markdown table.
    python tools/ncu_full_summary.py gpurun_out/r2r > profiles/r02_top_kernels_ncu_full.md"""
import csv
import subprocess
import sys
from pathlib import Path

METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__cluster_size', 'cluster size'), ('launch__registers_per_thread', 'registers / thread'),
    ('sm__cycles_elapsed.avg.per_second', 'SM clock'),
    ('sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'tcgen05 bf16 MMA op rate, % of peak'),
    ('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor pipe cycles active, %'),
    ('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor memory cycles active, %'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'shared-memory LSU wavefronts, % of peak'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/TEX throughput, %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput, %'), ('lts__t_bytes.sum', 'L2 bytes'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput, %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput, %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active, %'),
    ('smsp__inst_executed.sum', 'warp instructions executed'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy, %'),
]


def load(rep):
    out = subprocess.run(['ncu', '-i', str(rep), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
    d['_kernel'] = d.get('Kernel Name', ('?', ''))[0]
    return d


def main():
    root = Path(sys.argv[1])
    reps = sorted(root.glob('*.ncu-rep'))
    data = {r.stem: load(r) for r in reps}
    print('| metric | ' + ' | '.join(data) + ' |')
    print('|---|' + '---|' * len(data))
    print('| kernel | ' + ' | '.join(d['_kernel'][:60] for d in data.values()) + ' |')
    for key, label in METRICS:
        cells = []
        for d in data.values():
            v, u = d.get(key, ('n/a', ''))
            try:
                v = f'{float(v.replace(",", "")):.4g}'
            except ValueError:
                pass
            cells.append(f'{v} {u}'.strip())
        print(f'| {label} (`{key.split(".")[0] if len(key) < 60 else key.split(".")[-2][:40]}`) | ' + ' | '.join(cells) + ' |')


if __name__ == '__main__':
    main()
