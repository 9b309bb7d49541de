"""f2 micro-benchmark: one fused batch-assembly launch vs the reference's torch sequence (DataPreprocessor10.py:530-595) on the GPU."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import batch as OB            # the reference's op sequence, restated (used here only as the thing compared against)
from simple_rf_b200 import batch

dev = 'cuda'
t = {k: v.to(dev) for k, v in OB.synthetic_tables(num_views=3, h=756, w=1008, seed=0).items()}
indices, m_nerf, m_sd = (x.to(dev) for x in OB.synthetic_indices(t['pixel'].shape[0], 2048, 2048, seed=1))


def torch_sequence():
    n = indices.shape[0]
    pixel_id = -1 * torch.ones((n, 3), dtype=torch.int32).to(dev)
    target_rgb = -1 * torch.ones((n, 3)).to(dev)
    idx_nerf = indices[m_nerf]
    pixel_id[m_nerf] = t['pixel'][idx_nerf]
    target_rgb[m_nerf] = t['rgb'][idx_nerf]
    idx_sd = indices[m_sd]
    pixel_id[m_sd] = t['pixel'][idx_sd]
    d, e, p = -1 * torch.ones((n, 1)).to(dev), -1 * torch.ones((n, 1)).to(dev), -1 * torch.ones((n, 3)).to(dev)
    d[m_sd] = t['depth'][idx_sd]; e[m_sd] = t['error'][idx_sd]; p[m_sd] = t['points'][idx_sd]
    return pixel_id, target_rgb, d, e, p


def fused():
    return batch.assemble_batch(indices, m_sd, t['pixel'], t['rgb'], t['depth'], t['error'], t['points'])


for name, fn in (('reference torch sequence (CPU tensors created then moved, masked index_put)', torch_sequence), ('srf_assemble_batch', fused)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    print(f'{name}: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per 4096-ray batch (wall clock, launch-bound)')
