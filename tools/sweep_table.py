#!/usr/bin/env python
"""Markdown tables of `python tools/hbm_microbench.py --sweep` (one JSON line per case):  python tools/sweep_table.py sweep.jsonl"""
import json
import sys

rows = [json.loads(ln) for ln in open(sys.argv[1]) if ln.startswith('{') and 'kind' in ln]
TITLES = {'composite_fwd': 'composite forward (weights + per-ray maps)', 'composite_bwd': 'composite backward (hand-written)',
          'sample_pdf_merge': 'sample_pdf + merge (S_c = S, N_f = 2 S)'}
for kind, title in TITLES.items():
    print(f'\n## {title}\n')
    print('| rays \\ samples | 64 | 128 | 192 | 256 | 512 |')
    print('|---|---|---|---|---|---|')
    for e in (16, 18, 20, 22):
        cells = []
        for S in (64, 128, 192, 256, 512):
            hit = [r for r in rows if r['kind'] == kind and r['R'] == 1 << e and r['S'] == S]
            cells.append(f"{hit[0]['frac_of_hbm_peak']:.3f} ({hit[0]['ms']:.2f} ms, {hit[0]['GB/s']:.0f} GB/s)" if hit else 'n/a')
        print(f'| 2^{e} | ' + ' | '.join(cells) + ' |')
