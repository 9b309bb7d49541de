"""Fused patch-reprojection depth loss against the same computation in plain torch ops on the GPU
(the reference's formulation: 75 fancy-index gathers per patch set) and against the CPU oracle.  Device-timed."""
import json, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import generate_golden as GG
from oracle import losses as OL
from simple_rf_b200.loss_functions import patch_reprojection as PR

dev = 'cuda'


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def torch_reference_style(a, rule):
    """The reference's formulation with torch ops on the device (per-offset fancy-index gathers into preallocated patches)."""
    h, w = a['images'].shape[1:3]
    pid = a['pixel_id'].long()
    origins = a['poses'][:, :3, 3]
    dist = torch.sqrt(torch.sum(torch.square(origins[pid[:, 0]].unsqueeze(1).repeat([1, origins.shape[0], 1]) - origins), dim=2))
    vb = torch.kthvalue(dist, 2, dim=1)[1]
    poses_b = a['poses'][vb]
    permuter = torch.eye(3, device=dev); permuter[1:] *= -1

    def reproj(p):
        q = (a['k'][None] @ permuter[None] @ poses_b[:, :3, :3].transpose(1, 2) @ (p - poses_b[:, :3, 3])[..., None]).squeeze(-1)
        return (q[:, :2] / q[:, 2:]).round().long()
    p1 = reproj(a['rays_o'] + a['rays_d'] * a['depth1'][:, None])
    p2 = reproj(a['rays_o'] + a['rays_d'] * a['depth2'][:, None])
    xa, ya = pid[:, 1], pid[:, 2]
    val = lambda x, y: (x >= 2) & (x < w - 2) & (y >= 2) & (y < h - 2)
    va, v1, v2 = val(xa, ya), val(p1[:, 0], p1[:, 1]), val(p2[:, 0], p2[:, 1])
    x1, y1, x2, y2 = p1[:, 0].clip(0, w - 1), p1[:, 1].clip(0, h - 1), p2[:, 0].clip(0, w - 1), p2[:, 1].clip(0, h - 1)
    n = pid.shape[0]
    pa, pb1, pb2 = (torch.zeros(n, 5, 5, 3, device=dev) for _ in range(3))
    padded = torch.nn.functional.pad(a['images'], (0, 0, 0, 2, 0, 2))
    for i, oy in enumerate(range(-2, 3)):
        for j, ox in enumerate(range(-2, 3)):
            pa[:, i, j] = padded[pid[:, 0], ya + oy, xa + ox]
            pb1[:, i, j] = padded[vb, y1 + oy, x1 + ox]
            pb2[:, i, j] = padded[vb, y2 + oy, x2 + ox]
    r1 = torch.sqrt(torch.mean(torch.square(pa - pb1), dim=(1, 2, 3)))
    r2 = torch.sqrt(torch.mean(torch.square(pa - pb2), dim=(1, 2, 3)))
    m1 = ((r1 < r2) | ~v2) & (r1 < 0.1) & v1 & va
    m2 = ((r2 < r1) | ~v1) & (r2 < 0.1) & v2 & va
    return m1, m2


for n, (h, w) in ((2048, (378, 504)), (4096, (756, 1008)), (65536, (756, 1008))):
    c = GG.patch_loss_inputs(num_rays=n, h=h, w=w, seed=1)
    a = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in c.items()}
    fused = timed(lambda: PR.patch_reprojection_masks(a['rays_o'], a['rays_d'], a['depth1'], a['depth2'], a['pixel_id'], a['poses'], a['k'],
                                                      a['images'], (5, 5), 0.1, True))
    plain = timed(lambda: torch_reference_style(a, True), n=5)
    m1, m2 = PR.patch_reprojection_masks(a['rays_o'], a['rays_d'], a['depth1'], a['depth2'], a['pixel_id'], a['poses'], a['k'], a['images'],
                                         (5, 5), 0.1, True)
    t1, t2 = torch_reference_style(a, True)
    t0 = time.perf_counter()
    OL.patch_reprojection_masks(c['rays_o'], c['rays_d'], c['depth1'], c['depth2'], c['pixel_id'], c['poses'], c['k'], c['images'], (5, 5), 0.1, True)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({'case': f'patch_reprojection_masks N={n} frame={h}x{w}', 'fused_ms': round(fused, 4), 'torch_ops_on_gpu_ms': round(plain, 3),
                      'cpu_oracle_ms': round(cpu_ms, 1), 'speedup_vs_torch_gpu': round(plain / fused, 1),
                      'gathered_GB_per_s': round(n * 225 * 4 / fused / 1e6, 1),
                      'mask_agreement_with_torch_gpu': round(float(((m1 == t1) & (m2 == t2)).float().mean()), 6)}), flush=True)
