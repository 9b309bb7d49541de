"""GPU parity: volume-rendering compositing forward + hand-written backward (through the C ABI)."""
import pytest
import torch

from oracle import composite as C

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-3          # north_star: rgb/depth/weights within 1e-3 on fp32 paths (measured far below)


def _inputs(R, S, ndc, seed, empty_rays=True):
    g = torch.Generator().manual_seed(seed)
    sigma = torch.relu(torch.randn(R, S, generator=g)) * 10
    if empty_rays and R > 2:
        sigma[0] = 0            # empty ray: acc = 0, depth = 0/(1e-6)
        sigma[1] = 1e4          # opaque at the first sample: q hits the 1e-10 floor
    rgb = torch.rand(R, S, 3, generator=g)
    z = torch.sort(torch.rand(R, S, generator=g), -1)[0]
    if not ndc:
        z = 2 + 4 * z
    ro = torch.randn(R, 3, generator=g) * 0.1
    rd = torch.randn(R, 3, generator=g) * 0.3 - torch.tensor([0, 0, 1.])
    dn = torch.randn(R, 3, generator=g)
    return sigma, rgb, z, ro, rd, dn


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(1.0, b.double().abs().max().item())


@pytest.mark.parametrize('tag,ndc,white', [('ndc', True, False), ('world', False, True)])
def test_composite_vs_reference_golden(golden, tag, ndc, white):
    from simple_rf_b200 import ops
    g = golden('composite')
    a = {k[len(tag) + 1:]: v for k, v in g.items() if k.startswith(tag + '_')}
    d = lambda k: a[k].to(DEV)
    sigma = d('sigma').requires_grad_()
    rgb = d('rgb').requires_grad_()
    out = ops.composite(sigma, rgb, d('z'), d('rays_o'), d('rays_d'), d('rays_d_ndc'), ndc=ndc, white_bkgd=white)
    keys = ('rgb', 'acc', 'depth', 'depth_var', 'weights', 'alpha', 'visibility') + (('depth_ndc', 'depth_var_ndc') if ndc else ())
    for k in keys:
        assert _rel(out[k].detach().cpu(), a[f'out_{k}']) <= 2e-5, k
    loss = sum((out[k] * d(f'up_{k}')).sum() for k in ('rgb', 'acc', 'depth', 'depth_var', 'weights') + (('depth_ndc', 'depth_var_ndc') if ndc else ()))
    loss.backward()
    assert _rel(sigma.grad.cpu(), a['g_sigma']) <= 2e-4          # stated gradient tolerance (relative to max |g|)
    assert _rel(rgb.grad.cpu(), a['g_rgb']) <= 1e-5


@pytest.mark.parametrize('R,S,ndc,scale,white', [(1000, 64, True, 1.0, False), (257, 192, True, 1.0, True),
                                                 (64, 462, True, 25.0, False), (33, 1083, True, 25.0, False),
                                                 (100, 45, False, 1.0, False), (5, 1, True, 1.0, False),
                                                 (3, 31, False, 25.0, True), (0, 64, True, 1.0, False),
                                                 (7, 2, True, 1.0, False), (9, 3, False, 1.0, False), (130, 130, True, 1.0, False),
                                                 (41, 515, True, 25.0, True), (19, 128, False, 1.0, False), (17, 256, True, 1.0, False)])
def test_composite_forward_backward_vs_oracle(R, S, ndc, scale, white):
    from simple_rf_b200 import ops
    sigma, rgb, z, ro, rd, dn = _inputs(R, S, ndc, seed=R + S)
    ref = C.composite(sigma, rgb, z, ro, rd, dn, ndc=ndc, distance_scale=scale, white_bkgd=white)
    sg = sigma.to(DEV).requires_grad_()
    cg = rgb.to(DEV).requires_grad_()
    out = ops.composite(sg, cg, z.to(DEV), ro.to(DEV), rd.to(DEV), dn.to(DEV), ndc=ndc, white_bkgd=white, distance_scale=scale)
    if R == 0:
        assert out['weights'].shape == (0, S)
        return
    keys = ['rgb', 'acc', 'depth', 'weights', 'alpha', 'visibility'] + (['depth_ndc'] if ndc else [])
    for k in keys:
        assert (out[k].detach().cpu() - ref[k]).abs().max().item() <= TOL * max(1.0, ref[k].abs().max().item()), k
    for k in ['depth_var'] + (['depth_var_ndc'] if ndc else []):
        assert _rel(out[k].detach().cpu(), ref[k]) <= TOL, k
    g = torch.Generator().manual_seed(1)
    ups = {k: torch.rand(ref[k].shape, generator=g) for k in ['rgb', 'acc', 'depth', 'depth_var', 'weights'] + (['depth_ndc', 'depth_var_ndc'] if ndc else [])}
    loss = sum((out[k] * ups[k].to(DEV)).sum() for k in ups)
    loss.backward()
    dd = lambda t: t.double()
    gs, gc = C.composite_backward(dd(sigma), dd(rgb), dd(z), dd(ro), dd(rd), dd(dn), ndc=ndc, distance_scale=scale,
                                  white_bkgd=white, g_rgb=dd(ups['rgb']), g_acc=dd(ups['acc']), g_depth=dd(ups['depth']),
                                  g_depth_var=dd(ups['depth_var']), g_depth_ndc=dd(ups['depth_ndc']) if ndc else None,
                                  g_depth_var_ndc=dd(ups['depth_var_ndc']) if ndc else None, g_weights=dd(ups['weights']))
    assert _rel(sg.grad.cpu(), gs) <= 1e-3, 'g_sigma'      # relative to max |g| per tensor
    assert _rel(cg.grad.cpu(), gc) <= 1e-5, 'g_rgb'


@pytest.mark.parametrize('S', [64, 77, 192])
def test_composite_unaligned_views(S):
    """Per-sample tensors that do not start on a 16-byte boundary take the scalar load / store path: same results."""
    from simple_rf_b200 import ops
    R = 50
    sigma, rgb, z, ro, rd, dn = _inputs(R, S, True, seed=S)

    def shifted(t, k):
        buf = torch.zeros(t.numel() + k, device=DEV)
        buf[k:] = t.reshape(-1).to(DEV)
        v = buf[k:].view(t.shape)
        assert v.data_ptr() % 16 == (4 * k) % 16
        return v
    a = ops.composite(sigma.to(DEV).requires_grad_(), rgb.to(DEV).requires_grad_(), z.to(DEV), ro.to(DEV), rd.to(DEV), dn.to(DEV), ndc=True)
    sg, cg = shifted(sigma, 1).requires_grad_(), shifted(rgb, 3).requires_grad_()
    b = ops.composite(sg, cg, shifted(z, 2), ro.to(DEV), rd.to(DEV), dn.to(DEV), ndc=True)
    for k in ('rgb', 'acc', 'depth', 'depth_ndc', 'depth_var', 'weights', 'alpha', 'visibility'):
        assert torch.equal(a[k], b[k]), k
    ups = torch.rand(R, 3, device=DEV)
    (b['rgb'] * ups).sum().backward()
    ref = C.composite(sigma.requires_grad_(), rgb.requires_grad_(), z, ro, rd, dn, ndc=True)
    (ref['rgb'] * ups.cpu()).sum().backward()
    assert _rel(sg.grad.cpu(), sigma.grad) <= 1e-3 and _rel(cg.grad.cpu(), rgb.grad) <= 1e-5


def test_composite_without_rgb_and_inference_mode():
    from simple_rf_b200 import ops
    sigma, rgb, z, ro, rd, dn = _inputs(300, 192, True, seed=2)
    ref = C.composite(sigma, None, z, ro, rd, dn, ndc=True, distance_scale=25.0)
    with torch.no_grad():
        out = ops.composite(sigma.to(DEV), None, z.to(DEV), ro.to(DEV), rd.to(DEV), dn.to(DEV), ndc=True,
                            distance_scale=25.0, per_sample=False)
    assert 'rgb' not in out and 'alpha' not in out
    for k in ('acc', 'depth', 'depth_ndc', 'weights'):
        assert (out[k].cpu() - ref[k]).abs().max().item() <= TOL * max(1.0, ref[k].abs().max().item()), k


def test_composite_linearity_at_full_size():
    """Size-independent property at the microbench scale: rgb_map is linear in the per-sample colours and
    acc/weights do not depend on them."""
    from simple_rf_b200 import ops
    R, S = 1 << 16, 192
    g = torch.Generator(device=DEV).manual_seed(0)
    sigma = torch.relu(torch.randn(R, S, device=DEV, generator=g)) * 10
    c1 = torch.rand(R, S, 3, device=DEV, generator=g)
    c2 = torch.rand(R, S, 3, device=DEV, generator=g)
    z = torch.sort(torch.rand(R, S, device=DEV, generator=g), -1)[0]
    ro = torch.randn(R, 3, device=DEV, generator=g) * 0.1
    rd = torch.randn(R, 3, device=DEV, generator=g) * 0.3 - torch.tensor([0, 0, 1.], device=DEV)
    dn = torch.randn(R, 3, device=DEV, generator=g)
    with torch.no_grad():
        a = ops.composite(sigma, c1, z, ro, rd, dn, ndc=True)
        b = ops.composite(sigma, c2, z, ro, rd, dn, ndc=True)
        c = ops.composite(sigma, 0.25 * c1 + 0.75 * c2, z, ro, rd, dn, ndc=True)
    assert torch.equal(a['weights'], b['weights']) and torch.equal(a['acc'], c['acc'])
    assert (c['rgb'] - (0.25 * a['rgb'] + 0.75 * b['rgb'])).abs().max().item() <= 1e-5
    assert bool((a['acc'] <= 1 + 1e-4).all()) and bool((a['weights'] >= 0).all())
