"""GPU parity: ray generation, stratified depths, inverse-CDF resampling + merge (through the C ABI)."""
import pytest
import torch

from oracle import fixtures as FX
from oracle import rays as RY
from oracle import sampling as SP

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _tables(model_configs):
    from simple_rf_b200 import ops
    return ops.camera_tables(model_configs['intrinsics'], model_configs['extrinsics'], DEV)


@pytest.mark.parametrize('name,half,flip,vd_ndc', [('nerf', False, False, False), ('tensorf', True, True, True)])
def test_raygen_vs_oracle_and_golden(golden, golden_configs, name, half, flip, vd_ndc):
    from simple_rf_b200 import ops
    configs, mc = golden_configs(name)
    g = golden(f'{name}_eval')
    h, w = mc['resolution']
    pid = g['pixel_id'].to(DEV)
    out = ops.raygen(pid, _tables(mc), h, w, mc['near'], half_pixel=half, flip_x=flip, ndc=True, viewdirs_from_ndc=vd_ndc)
    for k, t in zip(('rays_o', 'rays_d', 'rays_o_ndc', 'rays_d_ndc', 'view_dirs'), out):
        # rays: <= 1e-6 abs (SURVEY.md §8c ledger); scaled by magnitude for the unnormalised arrays
        tol = 1e-6 * max(1.0, g[k].abs().max().item())
        assert (t.cpu() - g[k]).abs().max().item() <= tol, k
    # larger seeded batch, odd size, against the live oracle
    K = torch.tensor(mc['intrinsics']); E = torch.tensor(mc['extrinsics'])
    pid = FX.random_pixels(4099, K.shape[0], h, w, seed=9)
    ro, rd = RY.camera_rays(pid, K, E, half_pixel=half, flip_x=flip)
    img = pid[:, 0].long()
    on, dn = RY.ndc_rays(ro, rd, h, w, K[img, 0, 0], K[img, 1, 1], mc['near'])
    vd = RY.view_dirs(dn if vd_ndc else rd)
    out = ops.raygen(pid.to(DEV), _tables(mc), h, w, mc['near'], half_pixel=half, flip_x=flip, ndc=True, viewdirs_from_ndc=vd_ndc)
    for ref, t in zip((ro, rd, on, dn, vd), out):
        assert (t.cpu() - ref).abs().max().item() <= 1e-6 * max(1.0, ref.abs().max().item())


def test_raygen_empty_and_world():
    from simple_rf_b200 import ops
    K = torch.tensor([[[100., 0, 50], [0, 100., 40], [0, 0, 1]]]); E = torch.eye(4)[None]
    tabs = ops.camera_tables(K, E, DEV)
    out = ops.raygen(torch.zeros(0, 3, dtype=torch.int32, device=DEV), tabs, 80, 100, 1.0, half_pixel=False, flip_x=False,
                     ndc=False, viewdirs_from_ndc=False)
    assert out[0].shape == (0, 3) and out[2] is None
    pid = torch.tensor([[0, 50, 40], [0, 0, 0]], dtype=torch.int32)
    ro, rd = RY.camera_rays(pid, K, E, half_pixel=False, flip_x=False)
    out = ops.raygen(pid.to(DEV), tabs, 80, 100, 1.0, half_pixel=False, flip_x=False, ndc=False, viewdirs_from_ndc=False)
    assert torch.allclose(out[1].cpu(), rd, atol=1e-6) and torch.allclose(out[4].cpu(), RY.view_dirs(rd), atol=1e-6)


@pytest.mark.parametrize('S,R', [(64, 1000), (462, 37), (1, 5)])
def test_stratified_bit_exact(S, R):
    from simple_rf_b200 import ops
    ladder = SP.coarse_depths(S, 0.0, 1.0)
    torch.manual_seed(S + R)
    jitter = torch.rand(R, S)
    ref = SP.stratified_depths(ladder, R, jitter)
    got = ops.stratified_z(ladder.to(DEV), R, jitter.to(DEV))
    assert torch.equal(got.cpu(), ref)
    assert torch.equal(ops.stratified_z(ladder.to(DEV), R).cpu(), ladder.expand(R, S))
    fast = ops.stratified_z(ladder.to(DEV), R, philox_seed=7).cpu()
    if S > 1:
        mids = .5 * (ladder[1:] + ladder[:-1])
        lo = torch.cat([ladder[:1], mids]); hi = torch.cat([mids, ladder[-1:]])
        assert bool(((fast >= lo) & (fast <= hi)).all())
        assert fast.std(0).mean() > 0


def test_sample_pdf_bit_exact_vs_reference_golden(golden):
    from simple_rf_b200 import ops
    g = golden('sample_pdf')
    for tag in 'abc':
        z, w, u = g[f'{tag}_z'], g[f'{tag}_weights'], g[f'{tag}_u']
        N = u.shape[1]
        z_f, s, b, a = ops.sample_pdf_merge(z.to(DEV), w.to(DEV), N, u=u.to(DEV), return_indices=True)
        assert torch.equal(b.cpu(), g[f'{tag}_below']) and torch.equal(a.cpu(), g[f'{tag}_above']), tag
        assert torch.equal(s.cpu(), g[f'{tag}_samples']), tag
        assert torch.equal(z_f.cpu(), g[f'{tag}_z_fine']), tag


@pytest.mark.parametrize('S,N,R,det', [(64, 128, 20000, True), (64, 128, 20000, False), (3, 5, 11, False),
                                       (9, 1, 7, True), (192, 64, 300, False), (462, 128, 64, True),
                                       (1083, 128, 16, False)])
def test_sample_pdf_bit_exact_vs_oracle(S, N, R, det):
    """Indices, samples and merged depths are bit-identical to the ATen CPU path on the same inputs."""
    from simple_rf_b200 import ops
    g = torch.Generator().manual_seed(S * 7 + N)
    z = torch.sort(torch.rand(R, S, generator=g), -1)[0]
    w = torch.rand(R, S, generator=g) ** 5
    w[::3] *= 1e-5
    w[1::7, S // 2:] = 0
    if det:
        u_row = torch.linspace(0., 1., steps=N)
        u = u_row.expand(R, N).contiguous()
        u_dev = u_row.to(DEV)
    else:
        u = torch.rand(R, N, generator=g)
        u_dev = u.to(DEV)
    z_ref, s_ref, b_ref, a_ref = SP.fine_depths(z, w, u)
    z_f, s, b, a = ops.sample_pdf_merge(z.to(DEV), w.to(DEV), N, u=u_dev, return_indices=True)
    assert torch.equal(b.cpu(), b_ref) and torch.equal(a.cpu(), a_ref)
    assert torch.equal(s.cpu(), s_ref)
    assert torch.equal(z_f.cpu(), z_ref)
    # size-independent property: merged depths are sorted and contain the coarse depths
    assert bool((z_f[:, 1:] >= z_f[:, :-1]).all())


@pytest.mark.parametrize('det', [True, False])
def test_sample_pdf_extreme_weights_and_ties(det):
    """Weights far outside [0, 1] (fp64 prefix sums no longer exact: the sequential order must be reproduced), repeated
    coarse depths and zero-weight stretches (ties in the merge): still bit-identical to the ATen CPU path."""
    from simple_rf_b200 import ops
    R, S, N = 4000, 64, 128
    g = torch.Generator().manual_seed(5)
    z = torch.sort(torch.rand(R, S, generator=g), -1)[0]
    z[::5, 10:14] = z[::5, 10:11]                      # repeated depths
    w = torch.rand(R, S, generator=g)
    w[::2, 7] = 3e5                                     # one dominant bin: the smallest pdf value drops below 2^-28
    w[1::4, 20:40] = 0
    w[2::9] = 0                                         # all-zero rows: uniform pdf from the 1e-5 floor
    u = (torch.linspace(0., 1., steps=N).expand(R, N) if det else torch.rand(R, N, generator=g)).contiguous()
    z_ref, s_ref, b_ref, a_ref = SP.fine_depths(z, w, u)
    z_f, s_, b, a = ops.sample_pdf_merge(z.to(DEV), w.to(DEV), N, u=u.to(DEV), return_indices=True)
    assert torch.equal(b.cpu(), b_ref) and torch.equal(a.cpu(), a_ref)
    assert torch.equal(s_.cpu(), s_ref)
    assert torch.equal(z_f.cpu(), z_ref)


def test_sample_pdf_philox_mode_is_sorted_superset():
    from simple_rf_b200 import ops
    R, S, N = 513, 64, 128
    g = torch.Generator().manual_seed(0)
    z = torch.sort(torch.rand(R, S, generator=g), -1)[0].to(DEV)
    w = torch.rand(R, S, generator=g).to(DEV)
    z_f = ops.sample_pdf_merge(z, w, N, philox_seed=3)
    assert z_f.shape == (R, S + N) and bool((z_f[:, 1:] >= z_f[:, :-1]).all())
    assert bool((z_f.min(1)[0] == z[:, 0]).all())
