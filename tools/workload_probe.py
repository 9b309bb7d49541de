#!/usr/bin/env python
"""Run single bench.py workloads on one GPU and print their per-kernel CUDA-event tables (A/B of kernel variants selected by
environment variables, e.g. SRF_VM_BWD_VARIANT=0|4|8|16).   python tools/workload_probe.py tensorf_train nerf_train trajectory"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

torch.cuda.set_device(0)
dist = bench.Dist(0, 1, torch.device('cuda', 0))
for name in sys.argv[1:] or ['tensorf_train']:
    if name == 'tensorf_train':
        r = bench.train_workload('tensorf', dist, 'weak')
    elif name == 'nerf_train':
        r = bench.train_workload('nerf', dist, 'weak')
    elif name == 'trajectory':
        r = bench.tensorf_trajectory(dist)
    else:
        raise SystemExit(f'unknown workload {name}')
    print(json.dumps({'workload': name, 'ms_per_step': r['ms_per_step'], 'value': r['value'], 'unit': r['unit'],
                      'ms_per_step_eager': r.get('ms_per_step_eager'), 'cuda_graph': r.get('cuda_graph'), 'cuda_graph_error': r.get('cuda_graph_error'),
                      'kernels_ms_per_step': r['kernels_ms_per_step'],
                      'fracs': {k: round(v['frac'], 3) for k, v in r.get('rooflines', {}).items()}}), flush=True)
