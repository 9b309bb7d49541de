#!/usr/bin/env python
"""torch.profiler view of one training workload of bench.py: where the time of an iteration goes besides our own kernels
(torch elementwise / indexing kernels of the loss side, host launch gaps, synchronisations).
    python tools/profile_step.py nerf|tensorf [out.txt]"""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from simple_rf_b200 import synthetic  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else 'nerf'
out = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / 'gpurun_out' / f'profile_{kind}_train.txt'
torch.cuda.set_device(0)
dist = bench.Dist(0, 1, torch.device('cuda', 0))
captured = {}
orig = bench.timed_steps


def hook(dist_, fn, steps, warmup, collect=False):
    if 'done' not in captured:
        for s in range(4):
            fn(s)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for s in range(5):
                fn(4 + s)
            torch.cuda.synchronize()
        captured['done'] = prof
    return orig(dist_, fn, steps, warmup, collect)


bench.timed_steps = hook
r = bench.train_workload(kind, dist, 'weak', steps=5, warmup=2)
prof = captured['done']
text = prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=70)
text += '\n\n' + prof.key_averages().table(sort_by='cpu_time_total', row_limit=30, max_name_column_width=70)
text += f"\n\nms_per_step (unprofiled): {r['ms_per_step']:.3f}\n"
out.parent.mkdir(parents=True, exist_ok=True)
out.write_text(text)
print(text[-6000:])
