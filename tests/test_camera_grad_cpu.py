"""Learnable cameras, host-side pieces on the CPU (the kernels they sit next to are covered by tests/test_gpu_camera_grad.py):
the differentiable ray restatement used by the backward of `_RaysFromCameras` and the ray gradients of compositing, both against
autograd through the oracle's restatement of the reference (oracle/rays.py, oracle/composite.py)."""
import pytest
import torch

from oracle import composite as OC
from oracle import fixtures as FX
from oracle import rays as RY


def _cameras(golden_configs):
    configs, mc = golden_configs('nerf')
    return torch.tensor(mc['intrinsics']).float(), torch.tensor(mc['extrinsics']).float(), mc


@pytest.mark.parametrize('flip_x,half_pixel,ndc,from_ndc', [(False, False, True, False), (True, True, True, True), (False, False, False, False)])
def test_rays_from_cameras_matches_oracle(golden_configs, flip_x, half_pixel, ndc, from_ndc):
    from simple_rf_b200 import camera_grad as CG
    K, E, mc = _cameras(golden_configs)
    h, w = mc['resolution']
    pid = FX.random_pixels(257, K.shape[0], h, w, seed=3)
    ro, rd, on, dn, vd = CG.rays_from_cameras(E, pid, K, h, w, mc['near'], half_pixel=half_pixel, flip_x=flip_x, ndc=ndc,
                                              viewdirs_from_ndc=from_ndc)
    ro_ref, rd_ref = RY.camera_rays(pid, K, E, half_pixel=half_pixel, flip_x=flip_x)
    assert torch.equal(ro, ro_ref) and torch.allclose(rd, rd_ref, rtol=0, atol=1e-7)
    if ndc:
        img = pid[:, 0].long()
        on_ref, dn_ref = RY.ndc_rays(ro_ref, rd_ref, h, w, K[img, 0, 0], K[img, 1, 1], mc['near'])
        assert torch.allclose(on, on_ref, rtol=1e-6, atol=1e-6) and torch.allclose(dn, dn_ref, rtol=1e-6, atol=1e-6)
        assert torch.allclose(vd, RY.view_dirs(dn_ref if from_ndc else rd_ref), rtol=1e-6, atol=1e-6)
    else:
        assert on is None and dn is None
        assert torch.allclose(vd, RY.view_dirs(rd_ref), rtol=1e-6, atol=1e-6)


def test_rays_from_cameras_carries_gradient_to_the_pose_correction(golden_configs):
    """r, t of the drop-in's ExtrinsicsLearner (SimpleNeRF17.py:817-842) receive the gradient autograd derives through the reference formulas."""
    from simple_rf_b200 import camera_grad as CG
    from simple_rf_b200.models.SimpleNeRF91 import ExtrinsicsLearner
    K, E, mc = _cameras(golden_configs)
    h, w = mc['resolution']
    learner = ExtrinsicsLearner(E.numpy(), learn_rotation=True, learn_translation=True)
    g = torch.Generator().manual_seed(1)
    learner.r.data.copy_(torch.randn(learner.r.shape, generator=g) * 0.02)
    learner.t.data.copy_(torch.randn(learner.t.shape, generator=g) * 0.05)
    pid = FX.random_pixels(64, K.shape[0], h, w, seed=4)
    views = learner(torch.arange(learner.num_frames))
    outs = CG.rays_from_cameras(views, pid, K, h, w, mc['near'], half_pixel=False, flip_x=False, ndc=True, viewdirs_from_ndc=False)
    probes = [torch.randn(o.shape, generator=g) for o in outs]
    sum((o * p).sum() for o, p in zip(outs, probes)).backward()
    got_r, got_t = learner.r.grad.clone(), learner.t.grad.clone()
    # the same through the oracle's per-ray formulas on per-ray matrices, as the reference evaluates them (SimpleNeRF17.py:113-114, :170-176)
    learner.zero_grad()
    img = pid[:, 0].long()
    per_ray = learner(img)
    x = torch.cat([pid[:, 1:].float(), torch.ones(pid.shape[0], 1)], 1)
    dirs = (torch.linalg.inv(K)[img] @ x[:, :, None])[:, :, 0] * torch.tensor([1., -1., -1.])
    rd = (dirs[:, None, :] * per_ray[:, :3, :3]).sum(-1)
    ro = per_ray[:, :3, 3]
    on, dn = RY.ndc_rays(ro, rd, h, w, K[img, 0, 0], K[img, 1, 1], mc['near'])
    ref = (ro, rd, on, dn, RY.view_dirs(rd))
    sum((o * p).sum() for o, p in zip(ref, probes)).backward()
    assert torch.allclose(got_r, learner.r.grad, rtol=1e-4, atol=1e-5) and torch.allclose(got_t, learner.t.grad, rtol=1e-4, atol=1e-5)
    assert float(got_r.abs().max()) > 0 and float(got_t.abs().max()) > 0


@pytest.mark.parametrize('ndc,scale', [(True, 1.0), (False, 1.0), (False, 25.0)])
def test_composite_ray_gradients_match_autograd(ndc, scale):
    """ops._composite_ray_gradients from the g_sigma of the backward kernel (here: autograd's) == autograd through the rays."""
    from simple_rf_b200 import ops
    g = torch.Generator().manual_seed(7)
    R, S = 33, 24
    sigma = torch.relu(torch.randn(R, S, generator=g) * 3).double().requires_grad_()
    rgb = torch.rand(R, S, 3, generator=g).double()
    z = torch.sort(torch.rand(R, S, generator=g).double() * (0.98 if ndc else 4.0) + (0.0 if ndc else 2.0), dim=-1).values
    if not ndc:
        z.requires_grad_()               # world space: the box-march depths depend on the pose (SimpleTensoRF09.py:388-400)
    rays_o = (torch.randn(R, 3, generator=g).double() * 0.3 - torch.tensor([0., 0., 0.5])).requires_grad_()
    rays_d = (torch.randn(R, 3, generator=g).double() * 0.3 - torch.tensor([0., 0., 1.0])).requires_grad_()
    rays_dn = (torch.randn(R, 3, generator=g).double() * 0.5 + torch.tensor([0., 0., 1.5])).requires_grad_() if ndc else None
    out = OC.composite(sigma, rgb, z, rays_o, rays_d, rays_dn, ndc=ndc, distance_scale=scale)
    keys = ['rgb', 'acc', 'depth', 'depth_var'] + (['depth_ndc', 'depth_var_ndc'] if ndc else [])
    probes = {k: torch.randn(out[k].shape, generator=g).double() for k in keys}
    probes['weights'] = torch.randn(R, S, generator=g).double()
    loss = sum((out[k] * p).sum() for k, p in probes.items())
    wrt = [sigma, rays_o, rays_d] + ([rays_dn] if ndc else [z])
    grads = torch.autograd.grad(loss, wrt, allow_unused=True)
    g_sigma = grads[0]
    got = ops._composite_ray_gradients(sigma.detach(), z.detach(), out['visibility'].detach(), rays_o.detach(), rays_d.detach(),
                                       None if rays_dn is None else rays_dn.detach(), out['acc'].detach(), ndc, scale, g_sigma,
                                       g_depth=probes['depth'], g_depth_var=probes['depth_var'], want_z=not ndc)
    want = [grads[1], grads[2], grads[3] if ndc else None, None if ndc else grads[3]]
    assert ndc or float(want[3].abs().max()) > 0
    for name, a, b in zip(('rays_o', 'rays_d', 'rays_d_ndc', 'z'), got, want):
        if b is None:
            assert a is None or float(a.abs().max()) == 0.0, name
            continue
        if a is None:
            a = torch.zeros_like(b)
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-10), (name, float((a - b).abs().max()))


def test_world_space_tensorf_pose_gradient_decomposition(golden_configs):
    """Simple-TensoRF without NDC and with learnable cameras: the drop-in never differentiates the whole render; it adds up (i) the view-direction
    gradient of the colour branch, (ii) `ops._composite_ray_gradients` (|d|, and the sample depths), (iii) the entry depth of the box march
    (`camera_grad.box_entry_depth`) and (iv) rays -> view matrices (`camera_grad.rays_from_cameras`).  Here every kernel is replaced by the
    oracle stage it is tested against, and the sum is compared with autograd through the oracle's whole render (which
    tests/test_oracle_cpu.py pins to r.grad / t.grad of the unmodified reference)."""
    from oracle import sampling as SP
    from oracle import tensorf as TF
    from simple_rf_b200 import camera_grad as CG
    from simple_rf_b200 import ops
    configs, mc = golden_configs('tensorf_world')
    cfg = configs['model']['coarse_model']
    t = FX.tensorf_sets(configs, seed=23, with_alpha=False)['coarse_model']
    K, E = torch.tensor(mc['intrinsics']).float(), torch.tensor(mc['extrinsics']).float()
    h, w = mc['resolution']
    pid = FX.random_pixels(24, K.shape[0], h, w, seed=12)
    gen = torch.Generator().manual_seed(2)
    r0, t0 = torch.randn(K.shape[0], 3, generator=gen) * 0.01, torch.randn(K.shape[0], 3, generator=gen) * 0.02
    res = t['resolution'].long()
    step = torch.mean((t['bbox'][1] - t['bbox'][0]).float() / (res - 1)) * cfg['num_voxels_per_sample']
    S, near, far = t['num_samples'], mc['near'], mc['far']
    flags = dict(half_pixel=True, flip_x=True, ndc=False, viewdirs_from_ndc=False)
    kw = dict(ndc=False, distance_scale=cfg['distance_scale'], weight_threshold=cfg['ray_marching_weight_threshold'])
    keys = ('rgb', 'depth', 'depth_var', 'acc')
    probes = {k: torch.randn((24, 3) if k == 'rgb' else (24,), generator=gen) for k in keys}

    # ---- the whole render under autograd (oracle)
    r, tt = r0.clone().requires_grad_(), t0.clone().requires_grad_()
    ro, rd = RY.camera_rays(pid, K, RY.pose_correction(E, r, tt), half_pixel=True, flip_x=True)
    z = SP.box_march_depths(ro, rd, t['bbox'], near, far, step, S)
    pts = ro[:, None, :] + rd[:, None, :] * z[..., None]
    out = TF.tensor_forward(t['params'], t['bbox'], pts, z, ro, rd, None, RY.view_dirs(rd), **kw)
    sum((out[k] * probes[k]).sum() for k in keys).backward()
    want_r, want_t = r.grad.clone(), tt.grad.clone()
    assert float(out['surface_mask'].float().mean()) > 0.01

    # ---- the drop-in's decomposition
    r, tt = r0.clone().requires_grad_(), t0.clone().requires_grad_()
    rays_o, rays_d, _, _, view_dirs = CG.rays_from_cameras(RY.pose_correction(E, r, tt), pid, K, h, w, near, **flags)      # (iv)
    ro_l, rd_l, vd_l = (x.detach().requires_grad_() for x in (rays_o, rays_d, view_dirs))
    z_values = SP.box_march_depths(ro_l.detach(), rd_l.detach(), t['bbox'], near, far, step, S)                             # srf_box_march_z
    entry = CG.box_entry_depth(ro_l, rd_l, t['bbox'].tolist(), near, far)                                                  # (iii)
    z_attached = z_values + (entry - entry.detach())[:, None]
    assert torch.equal(z_attached.detach(), z_values)
    pts = ro_l.detach()[:, None, :] + rd_l.detach()[:, None, :] * z_values[..., None]                                       # kernels see raw values
    mask = TF.validity_mask(pts, t['bbox'])
    pn = TF.normalize(pts, t['bbox'])
    sigma = TF.density(t['params'], pn, mask)[..., 0].detach().requires_grad_()
    with torch.no_grad():
        surface = OC.composite(sigma, None, z_values, ro_l, rd_l, None, ndc=False, distance_scale=cfg['distance_scale'])['weights'] > kw['weight_threshold']
    rgb = TF.vm_color(t['params'], pn, surface, vd_l)                                                                      # colour branch, (i)
    rgb_l = rgb.detach().requires_grad_()
    vr = OC.composite(sigma, rgb_l, z_values, ro_l.detach(), rd_l.detach(), None, ndc=False, distance_scale=cfg['distance_scale'])
    g_sigma, g_rgb = torch.autograd.grad(sum((vr[k] * probes[k]).sum() for k in keys), [sigma, rgb_l])                     # srf_composite_bwd
    g_o, g_d, _, g_z = ops._composite_ray_gradients(sigma.detach(), z_values, vr['visibility'].detach(), ro_l.detach(), rd_l.detach(), None,
                                                    vr['acc'].detach(), False, cfg['distance_scale'], g_sigma,
                                                    g_depth=probes['depth'], g_depth_var=probes['depth_var'], want_z=True)  # (ii)
    torch.autograd.backward([rgb, z_attached], [g_rgb, g_z])                 # -> vd_l.grad, and ro_l.grad / rd_l.grad through the entry depth
    g_rays_o = ro_l.grad + (g_o if g_o is not None else 0)
    g_rays_d = rd_l.grad + g_d
    torch.autograd.backward([rays_o, rays_d, view_dirs], [g_rays_o, g_rays_d, vd_l.grad])
    for name, got, want in (('r', r.grad, want_r), ('t', tt.grad, want_t)):
        rel = float((got - want).norm() / want.norm())
        assert rel <= 1e-3, (name, rel)
