"""Device-timed HBM-roofline microbenchmarks (BASELINE.json configs[4]): compositing fwd/bwd, sample_pdf+merge,
ray generation, TensoRF mask/compaction/gathers.  Algorithmic bytes per SURVEY.md §8d.  Prints one JSON line per case."""
import json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simple_rf_b200 import ops, tensorf_ops as T, _lib
from oracle import tensorf as TF

dev = 'cuda'
PEAK = json.loads((ROOT / 'MEASURED_PEAKS.json').read_text())['hbm_gbs'] if (ROOT / 'MEASURED_PEAKS.json').exists() else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=5):
    """Median device time of fn() with a cold L2.  The GPU first evicts L2 and then spins for ~1 ms (torch.cuda._sleep) while the host
    enqueues the start event, fn's launches and the stop event: without that head start the interval between the two events of a
    sub-100-us launch is the HOST time of the Python wrapper (allocations, ctypes call), not the kernel's."""
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        flush.zero_()                       # evict L2 (256 MB > 126 MB)
        torch.cuda._sleep(2_000_000)        # ~1 ms of GPU spinning: the launches below are queued before it ends
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def report(name, ms, nbytes, **extra):
    gbs = nbytes / ms / 1e6
    print(json.dumps({'case': name, 'ms': round(ms, 4), 'algorithmic_GB': round(nbytes / 1e9, 4), 'GB/s': round(gbs, 1),
                      'frac_of_hbm_peak': round(gbs / PEAK, 3), **extra}), flush=True)


g = torch.Generator(device=dev).manual_seed(0)


def sweep():
    """BASELINE.json configs[4] in full: R in 2^16..2^22 x S in {64,128,192,256,512}; compositing forward (weights only),
    backward and sample_pdf + merge (S_c = S, N_f = 2 S, random u in-kernel).  Inputs as SURVEY.md §8d prescribes."""
    torch.manual_seed(0)
    for R in (1 << 16, 1 << 18, 1 << 20, 1 << 22):
        for S in (64, 128, 192, 256, 512):
            sigma = torch.randn(R, S, device=dev, generator=g).relu_().mul_(10).requires_grad_()
            rgb = torch.rand(R, S, 3, device=dev, generator=g).requires_grad_()
            z = torch.rand(R, S, device=dev, generator=g)
            z = torch.sort(z, -1)[0]
            ro = torch.randn(R, 3, device=dev, generator=g) * .1
            rd = torch.nn.functional.normalize(torch.randn(R, 3, device=dev, generator=g), dim=-1) * (0.5 + torch.rand(R, 1, device=dev, generator=g))
            dn = torch.randn(R, 3, device=dev, generator=g)
            n = 3 if R * S >= (1 << 28) else 5
            with torch.no_grad():
                ms = timed(lambda: ops.composite(sigma, rgb, z, ro, rd, dn, ndc=True, per_sample=False), n)
            report(f'sweep composite_fwd R=2^{R.bit_length() - 1} S={S}', ms, R * S * 24 + R * 68, R=R, S=S, kind='composite_fwd')
            out = ops.composite(sigma, rgb, z, ro, rd, dn, ndc=True)
            keys = ('rgb', 'depth', 'depth_ndc', 'acc')
            loss_grads = [torch.rand_like(out[k]) for k in keys]
            ms = timed(lambda: torch.autograd.grad([out[k] for k in keys], [sigma, rgb], loss_grads, retain_graph=True), n)
            report(f'sweep composite_bwd R=2^{R.bit_length() - 1} S={S}', ms, R * S * 40 + R * 60, R=R, S=S, kind='composite_bwd')
            w = out['weights'].detach()
            del out, loss_grads
            try:
                ms = timed(lambda: ops.sample_pdf_merge(z, w, 2 * S, philox_seed=1), n)
                report(f'sweep sample_pdf_merge R=2^{R.bit_length() - 1} {S}->+{2 * S}', ms, R * (4 * (S - 2) + 4 * S + 4 * 3 * S), R=R, S=S,
                       kind='sample_pdf_merge')
            except Exception as e:
                print(json.dumps({'case': f'sweep sample_pdf_merge R=2^{R.bit_length() - 1} {S}->+{2 * S}', 'skipped': str(e)[:120]}), flush=True)
            del sigma, rgb, z, w, ro, rd, dn
            torch.cuda.empty_cache()


if '--sweep' in sys.argv:
    sweep()
    sys.exit(0)
sizes = [(1 << 18, 64), (1 << 18, 192), (1 << 20, 64), (1 << 20, 192), (1 << 18, 512)]
if '--probe' in sys.argv:          # one shape per run (ncu groups launches by kernel + grid): --probe [S]
    sizes = [(1 << 20, int(sys.argv[sys.argv.index('--probe') + 1]) if len(sys.argv) > sys.argv.index('--probe') + 1 else 192)]
if '--big' in sys.argv:
    sizes += [(1 << 22, 64), (1 << 22, 192), (1 << 20, 512)]
for R, S in sizes:
    sigma = (torch.relu(torch.randn(R, S, device=dev, generator=g)) * 10).requires_grad_()
    rgb = torch.rand(R, S, 3, device=dev, generator=g).requires_grad_()
    z = torch.sort(torch.rand(R, S, device=dev, generator=g), -1)[0]
    ro = torch.randn(R, 3, device=dev, generator=g) * .1
    rd = torch.randn(R, 3, device=dev, generator=g) * .3 - torch.tensor([0, 0, 1.], device=dev)
    dn = torch.randn(R, 3, device=dev, generator=g)
    with torch.no_grad():
        ms = timed(lambda: ops.composite(sigma, rgb, z, ro, rd, dn, ndc=True, per_sample=False))
    report(f'composite_fwd R={R} S={S}', ms, R * S * 24 + R * 68)
    with torch.no_grad():
        ms = timed(lambda: ops.composite(sigma, rgb, z, ro, rd, dn, ndc=True, per_sample=True))
    report(f'composite_fwd+alpha,vis R={R} S={S}', ms, R * S * 32 + R * 68)
    out = ops.composite(sigma, rgb, z, ro, rd, dn, ndc=True)
    loss_grads = [torch.rand_like(out[k]) for k in ('rgb', 'depth', 'depth_ndc', 'acc')]

    def bwd():
        torch.autograd.grad([out['rgb'], out['depth'], out['depth_ndc'], out['acc']], [sigma, rgb], loss_grads, retain_graph=True)
    ms = timed(bwd)
    report(f'composite_bwd R={R} S={S}', ms, R * S * 40 + R * 60)
    del out
    if S == 64 and '--probe' not in sys.argv:
        w = torch.rand(R, S, device=dev, generator=g)
        u = torch.rand(R, 128, device=dev, generator=g)
        ms = timed(lambda: ops.sample_pdf_merge(z, w, 128, u=u))
        report(f'sample_pdf_merge R={R} 64->+128 (u given)', ms, R * (4 * 62 + 4 * 64 + 4 * 128 + 4 * 192))
        u_det = torch.linspace(0., 1., steps=128, device=dev)
        ms = timed(lambda: ops.sample_pdf_merge(z, w, 128, u=u_det))
        report(f'sample_pdf_merge R={R} 64->+128 (deterministic u, test-time path)', ms, R * (4 * 62 + 4 * 64 + 4 * 192))
        ms = timed(lambda: ops.sample_pdf_merge(z, w, 128, philox_seed=1))
        report(f'sample_pdf_merge R={R} 64->+128 (philox)', ms, R * (4 * 62 + 4 * 64 + 4 * 192))
    del sigma, rgb, z

if '--probe' in sys.argv:
    sys.exit(0)
# ray generation
K = torch.tensor([[[815.13, 0, 504.], [0, 815.13, 378.], [0, 0, 1.]]]); E = torch.eye(4)[None]
tabs = ops.camera_tables(K, E, dev)
R = 1 << 22
pid = torch.stack([torch.zeros(R, dtype=torch.int32), torch.randint(0, 1008, (R,), dtype=torch.int32),
                   torch.randint(0, 756, (R,), dtype=torch.int32)], 1).to(dev)
ms = timed(lambda: ops.raygen(pid, tabs, 756, 1008, 1.0, half_pixel=False, flip_x=False, ndc=True, viewdirs_from_ndc=False))
report(f'raygen R={R}', ms, R * 72)

# TensoRF: 300^3-class grid (331x368x220, 1083 samples/ray), 4096-ray chunk, ~random occupancy
res = torch.tensor([331, 368, 220])
bbox = torch.tensor([[-1.5, -1.67, -1.0], [1.5, 1.67, 1.0]])
params = {k: v.to(dev) for k, v in TF.init_vm_params(res, [16, 4, 4], [48, 12, 12], generator=torch.Generator().manual_seed(0)).items()}
R, S = 4096, 1083
o = torch.cat([torch.rand(R, 2) * 2 - 1, -torch.ones(R, 1)], 1).to(dev) * torch.tensor([1.2, 1.3, 1.0], device=dev)
d = torch.cat([torch.randn(R, 2) * .1, 2 * torch.ones(R, 1)], 1).to(dev)
z = torch.sort(torch.rand(R, S, device=dev), -1)[0]
vol = (torch.rand(190, 190, 190) < 0.05).float().to(dev)
alpha = {'bits': T.pack_alpha_bits(vol), 'res': [190, 190, 190], 'box_min': bbox[0].tolist(), 'box_size': (bbox[1] - bbox[0]).tolist()}
ms = timed(lambda: T.validity_compact(o, d, z, bbox, alpha))
comp = T.validity_compact(o, d, z, bbox, alpha)
n = int(comp.count.item())
report(f'tensorf mask+compaction R={R} S={S} (valid {n / (R * S):.3f})', ms, R * S * (4 + 1) + n * 4)
comp_all = T.validity_compact(o, d, z, bbox, None)
n_all = int(comp_all.count.item())
geom = T.VmGeometry(o, d, z, bbox[0], bbox[1] - bbox[0], res)
planes = [params[f'matrices_density.{i}'] for i in range(3)]
lines = [params[f'vectors_density.{i}'] for i in range(3)]
pc, lc = T.to_channels_last(planes, lines)
chans = T._i3([16, 4, 4])
sigma = torch.zeros(R, S, 1, device=dev); feat = torch.empty(R * S, device=dev)
for name, c, nn in (('alpha-masked', comp, n), ('all in-box', comp_all, n_all)):
    ms = timed(lambda: _lib.call('srf_vm_density_fwd', *geom.args(c), T._ptrs(pc), T._ptrs(lc), chans, geom.res, 0, 0.0,
                                 _lib.ptr(sigma), _lib.ptr(feat), _lib.stream_handle()))
    report(f'vm_density_fwd {name} n={nn}', ms, nn * 576, unit_note='requested texel bytes (L2-level), not HBM', hbm_bytes=nn * 12)
gs = torch.rand(R, S, 1, device=dev)
gp = [torch.zeros_like(p) for p in pc]; gl = [torch.zeros_like(l) for l in lc]
ms = timed(lambda: _lib.call('srf_vm_density_bwd', *geom.args(comp_all), T._ptrs(pc), T._ptrs(lc), chans, geom.res, 0, 0.0, _lib.ptr(gs),
                             _lib.ptr(feat), T._ptrs(gp), T._ptrs(gl), _lib.stream_handle()))
report(f'vm_density_bwd all in-box n={n_all}', ms, n_all * 576, unit_note='requested texel bytes scattered (L2 atomics)')
wts = torch.rand(R, S, device=dev) ** 12
surf = T.threshold_compact(wts, 1e-4)
ns = int(surf.count.item())
cplanes = [params[f'matrices_color.{i}'] for i in range(3)]; clines = [params[f'vectors_color.{i}'] for i in range(3)]
vd = torch.nn.functional.normalize(torch.randn(R, 3, device=dev), dim=-1)
with torch.no_grad():
    ms = timed(lambda: T.vm_color_rows(geom, surf, vd, cplanes, clines))
report(f'vm_color_rows n={ns} ({ns / (R * S):.3f} of samples)', ms, ns * 1728, unit_note='requested texel bytes; includes channels-last cache rebuild')
