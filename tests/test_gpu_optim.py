"""GPU parity of the fused optimiser tail (f4): srf_adam_step through the C ABI against the CPU oracle, and the
FusedFlatAdam wrapper against torch.optim.Adam (same trajectory, same state_dict format, parameter-set changes)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def test_adam_kernel_vs_oracle():
    from simple_rf_b200 import _lib as L
    from oracle import adam as OA
    g_ = torch.Generator().manual_seed(1)
    for n in (1, 3, 4, 1023, 262144 + 5):
        p = torch.randn(n, generator=g_); g = torch.randn(n, generator=g_) * 1e-2
        m = torch.randn(n, generator=g_) * 1e-2; v = torch.rand(n, generator=g_) * 1e-4
        for step, lr, wd in ((1, 5e-4, 0.0), (7, 2e-2, 0.0), (1000, 1e-3, 1e-2)):
            dp, dg, dm, dv = (x.clone().to(DEV) for x in (p, g, m, v))
            L.call('srf_adam_step', L.ptr(dp), L.ptr(dg), L.ptr(dm), L.ptr(dv), n, lr, 0.9, 0.999, 1e-8, wd, step, L.stream_handle())
            rp, rm, rv = OA.adam_step(p.numpy(), g.numpy(), m.numpy(), v.numpy(), step, lr, weight_decay=wd)
            assert np.abs(dp.cpu().numpy() - rp).max() <= 1e-6 * max(1.0, np.abs(rp).max())
            assert np.abs(dm.cpu().numpy() - rm).max() <= 1e-6 * np.abs(rm).max() and np.abs(dv.cpu().numpy() - rv).max() <= 1e-6 * np.abs(rv).max()


def _make(seed):
    torch.manual_seed(seed)
    shapes = [(256, 63), (256,), (1, 16, 33, 21), (3, 128), (3,)]
    params = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    params[1].requires_grad_(False)                    # frozen parameter inside a group
    opt = torch.optim.Adam([{'params': params[:3], 'lr': 5e-4, 'name': 'a'}, {'params': params[3:], 'lr': 2e-2, 'name': 'b'}],
                           betas=(0.9, 0.999))
    return params, opt


def test_fused_adam_follows_torch():
    from simple_rf_b200 import optim
    pa, oa = _make(0)
    pb, ob = _make(0)
    assert optim.supports(ob)
    fused = optim.FusedFlatAdam(ob)
    g_ = torch.Generator(device=DEV).manual_seed(5)
    for it in range(12):
        grads = [torch.randn(p.shape, device=DEV, generator=g_) * 0.1 for p in pa]
        for o in (oa, ob):
            o.zero_grad(set_to_none=True)
        for i, (a, b) in enumerate(zip(pa, pb)):
            if not a.requires_grad or (i == 4 and it % 3 == 0):          # a parameter that sometimes has no gradient
                continue
            a.grad = grads[i].clone(); b.grad = grads[i].clone()
        for o in (oa, ob):                                                # the trainer rescales lr every iteration
            for gr in o.param_groups:
                gr['lr'] *= 0.999
        versions = [b._version for b in pb]
        oa.step(); ob.step()
        assert all(b._version > v for b, v in zip(pb, versions) if b.grad is not None)      # derived caches key on _version
        for a, b in zip(pa, pb):
            assert (a - b).abs().max().item() <= 2e-6 * max(1.0, a.abs().max().item()), it
    import copy
    sa, sb = copy.deepcopy(oa.state_dict()), copy.deepcopy(ob.state_dict())   # as a checkpoint round trip would (no aliasing)
    assert sa['param_groups'][0]['params'] == sb['param_groups'][0]['params']
    for k in sa['state']:
        assert float(sa['state'][k]['step']) == float(sb['state'][k]['step'])
        for key in ('exp_avg', 'exp_avg_sq'):
            assert (sa['state'][k][key] - sb['state'][k][key]).abs().max().item() <= 1e-4 * sa['state'][k][key].abs().max().item()   # 12 steps of FMA-vs-separate rounding
    # resume: a fresh torch Adam loaded from the fused optimiser's state_dict continues identically, and the fused one
    # survives load_state_dict (its flat buffers are rebuilt from the loaded states)
    pc, oc = _make(0)
    with torch.no_grad():
        for c, b in zip(pc, pb):
            c.copy_(b)
    oc.load_state_dict(sb)
    ob.load_state_dict(sa)
    with torch.no_grad():
        for a, b in zip(pa, pb):
            b.copy_(a)
    for it in range(3):
        grads = [torch.randn(p.shape, device=DEV, generator=g_) * 0.1 for p in pa]
        for ps in (pa, pb, pc):
            for i, p in enumerate(ps):
                p.grad = grads[i].clone() if p.requires_grad else None
        oa.step(); ob.step(); oc.step()
    for a, b, c in zip(pa, pb, pc):
        assert (a - b).abs().max().item() <= 3e-6 * max(1.0, a.abs().max().item())
        assert (b - c).abs().max().item() <= 3e-5 * max(1.0, a.abs().max().item())
    fused.detach()
    assert ob.step.__func__ is torch.optim.Adam.step or ob.step is not None


def test_fused_adam_parameter_set_change():
    """TensoRF replaces its plane parameters and re-adds groups on the same optimiser (SimpleTensoRF09.py:916-944)."""
    from simple_rf_b200 import optim
    pa, oa = _make(1)
    pb, ob = _make(1)
    optim.FusedFlatAdam(ob)
    for o, ps in ((oa, pa), (ob, pb)):
        for p in ps:
            p.grad = torch.ones_like(p) if p.requires_grad else None
        o.step()
    new_a = torch.nn.Parameter(torch.full((5, 7), 0.5, device=DEV))
    new_b = torch.nn.Parameter(torch.full((5, 7), 0.5, device=DEV))
    for o, ps, new in ((oa, pa, new_a), (ob, pb, new_b)):
        del o.param_groups[1]
        for p in ps[3:]:
            o.state.pop(p, None)
        o.add_param_group({'params': [new], 'lr': 1e-2, 'name': 'b'})
        for p in ps[:3] + [new]:
            p.grad = torch.ones_like(p) * 0.3 if p.requires_grad else None
        o.step()
    assert (new_a - new_b).abs().max().item() <= 1e-6
    for a, b in zip(pa[:3], pb[:3]):
        assert (a - b).abs().max().item() <= 2e-6 * max(1.0, a.abs().max().item())
