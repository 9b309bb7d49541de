"""L2-resident read bandwidth of this B200 (denominator for the TensoRF gathers, which are reported in requested texel bytes per
second: SURVEY.md §8d asks for them "against measured L2 bandwidth").  A 48 MB fp32 buffer (< 126 MB L2, the size of the
appearance planes at the 300^3 grid) is summed 64 times inside ONE launch of torch's reduction kernel (stride-0 expansion), so all but
the first pass hit the L2.  Prints one JSON line."""
import json
import torch

dev = 'cuda'
out = {}
REPS = 64
for mb in (24, 48, 96, 512):
    x = torch.rand(mb * (1 << 20) // 4, device=dev)
    xe = x.expand(REPS, x.numel())          # one launch reads the same buffer REPS times (stride-0 outer dimension)
    for _ in range(2):
        xe.sum(dim=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    xe.sum(dim=1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out[f'{mb}MB'] = round(REPS * x.numel() * 4 / ms / 1e6, 1)
print(json.dumps({'case': 'L2-resident read bandwidth (torch sum, GB/s); 512 MB = HBM for comparison', **out}))
