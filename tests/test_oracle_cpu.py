"""CPU suite: the oracle against the committed golden fixtures (outputs of the unmodified reference),
and the numpy restatement of the ATen evaluation order against the torch ops themselves."""
import numpy as np
import pytest
import torch

from oracle import aten_order as AO
from oracle import composite as C
from oracle import fixtures as FX
from oracle import pipeline as P
from oracle import sampling as SP
from oracle import tensorf as TF


def _close(a, b, tol):
    return (a.double() - b.double()).abs().max().item() <= tol


@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_nerf_pipeline_matches_reference_golden(golden, golden_configs, mode):
    g = golden(f'nerf_{mode}')
    configs, model_configs = golden_configs('nerf')
    sets = FX.nerf_param_sets(configs, seed=int(g['param_seed']))
    torch.manual_seed(int(g['rng_seed']))
    with torch.no_grad():
        out = P.nerf_render_chunk(sets, configs, model_configs, g['pixel_id'], training=(mode == 'train'))
    # sampling stages do not depend on BLAS: bit-exact everywhere
    for k in ('rays_o', 'rays_d', 'z_vals_coarse'):
        assert torch.equal(out[k], g[k]), k
    # anything behind the MLP depends on the host's sgemm blocking: fp32 tolerance
    for k in g:
        if k in out and g[k].dtype == torch.float32:
            assert _close(out[k], g[k], 2e-4 * max(1.0, g[k].abs().max().item())), k


@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_nerf_variant_pipeline_matches_reference_golden(golden, golden_configs, mode):
    """World-space sampling (`ndc = False`), depths linear in disparity, white background (oracle/generate_golden.py::nerf_variant_configs)."""
    g = golden(f'nerf_variant_{mode}')
    configs, model_configs = golden_configs('nerf_variant')
    assert configs['data_loader']['ndc'] is False and configs['model']['lindisp'] and configs['model']['white_bkgd']
    sets = FX.nerf_param_sets(configs, seed=int(g['param_seed']))
    torch.manual_seed(int(g['rng_seed']))
    with torch.no_grad():
        out = P.nerf_render_chunk(sets, configs, model_configs, g['pixel_id'], training=(mode == 'train'))
    for k in ('rays_o', 'rays_d', 'z_vals_coarse'):
        assert torch.equal(out[k], g[k]), k
    assert not any('ndc' in k for k in out)
    for k in g:
        if k in out and g[k].dtype == torch.float32:
            assert _close(out[k], g[k], 2e-4 * max(1.0, g[k].abs().max().item())), k


def test_sample_pdf_matches_reference_golden(golden):
    g = golden('sample_pdf')
    for tag in 'abc':
        z_f, samples, below, above = SP.fine_depths(g[f'{tag}_z'], g[f'{tag}_weights'], g[f'{tag}_u'])
        assert torch.equal(samples, g[f'{tag}_samples'])
        assert torch.equal(z_f, g[f'{tag}_z_fine'])
        assert torch.equal(below, g[f'{tag}_below']) and torch.equal(above, g[f'{tag}_above'])


@pytest.mark.parametrize('tag,ndc,white', [('ndc', True, False), ('world', False, True)])
def test_composite_matches_reference_golden(golden, tag, ndc, white):
    g = golden('composite')
    a = {k[len(tag) + 1:]: v for k, v in g.items() if k.startswith(tag + '_')}
    out = C.composite(a['sigma'], a['rgb'], a['z'], a['rays_o'], a['rays_d'], a['rays_d_ndc'], ndc=ndc, white_bkgd=white)
    for k in ('rgb', 'acc', 'depth', 'depth_var', 'weights', 'alpha', 'visibility'):
        assert _close(out[k], a[f'out_{k}'], 1e-6 * max(1.0, a[f'out_{k}'].abs().max().item())), k
    dd = lambda t: None if t is None else t.double()
    gs, gc = C.composite_backward(
        dd(a['sigma']), dd(a['rgb']), dd(a['z']), dd(a['rays_o']), dd(a['rays_d']), dd(a['rays_d_ndc']), ndc=ndc,
        white_bkgd=white, g_rgb=dd(a['up_rgb']), g_acc=dd(a['up_acc']), g_depth=dd(a['up_depth']),
        g_depth_var=dd(a['up_depth_var']), g_depth_ndc=dd(a.get('up_depth_ndc')),
        g_depth_var_ndc=dd(a.get('up_depth_var_ndc')), g_weights=dd(a['up_weights']))
    scale = a['g_sigma'].abs().max().item()
    assert _close(gs / scale, a['g_sigma'] / scale, 3e-4)
    assert _close(gc, a['g_rgb'], 1e-5)


def test_composite_backward_closed_form_vs_autograd_fp64():
    g = torch.Generator().manual_seed(0)
    R, S = 7, 33
    sigma = (torch.relu(torch.randn(R, S, generator=g, dtype=torch.float64)) * 5).requires_grad_()
    rgb = torch.rand(R, S, 3, generator=g, dtype=torch.float64).requires_grad_()
    z = torch.sort(torch.rand(R, S, generator=g, dtype=torch.float64), -1)[0]
    ro = torch.randn(R, 3, generator=g, dtype=torch.float64) * .1
    rd = torch.randn(R, 3, generator=g, dtype=torch.float64) * .3 - torch.tensor([0, 0, 1.], dtype=torch.float64)
    dn = torch.randn(R, 3, generator=g, dtype=torch.float64)
    out = C.composite(sigma, rgb, z, ro, rd, dn, ndc=True, distance_scale=25.0, white_bkgd=True)
    ups = {k: torch.rand(out[k].shape, generator=g, dtype=torch.float64)
           for k in ('rgb', 'acc', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc', 'weights')}
    loss = sum((out[k] * ups[k]).sum() for k in ups)
    gs, gc = torch.autograd.grad(loss, [sigma, rgb])
    gs2, gc2 = C.composite_backward(sigma.detach(), rgb.detach(), z, ro, rd, dn, ndc=True, distance_scale=25.0,
                                    white_bkgd=True, g_rgb=ups['rgb'], g_acc=ups['acc'], g_depth=ups['depth'],
                                    g_depth_var=ups['depth_var'], g_depth_ndc=ups['depth_ndc'],
                                    g_depth_var_ndc=ups['depth_var_ndc'], g_weights=ups['weights'])
    assert (gs - gs2).abs().max() <= 1e-9 * gs.abs().max()
    assert (gc - gc2).abs().max() <= 1e-12


@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_tensorf_pipeline_matches_reference_golden(golden, golden_configs, mode):
    g = golden(f'tensorf_{mode}')
    configs, model_configs = golden_configs('tensorf')
    sets = FX.tensorf_sets(configs, seed=int(g['param_seed']), with_alpha=bool(g['with_alpha']))
    torch.manual_seed(int(g['rng_seed']))
    with torch.no_grad():
        out = P.tensorf_render_chunk(sets, configs, model_configs, g['pixel_id'], training=(mode == 'train'))
    for k in ('rays_o', 'rays_d', 'rays_o_ndc', 'rays_d_ndc', 'view_dirs', 'z_vals_coarse'):
        assert torch.equal(out[k], g[k]), k
    assert torch.equal(out['validity_mask_coarse'], g['validity_mask_coarse'])      # bit-exact contract
    for k in g:
        if k in out and g[k].dtype == torch.float32:
            assert _close(out[k], g[k], 2e-4 * max(1.0, g[k].abs().max().item())), k


@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_tensorf_world_space_pipeline_matches_reference_golden(golden, golden_configs, mode):
    """`data_loader.ndc = False`: box-march depths (SimpleTensoRF09.py:388-400), world-space points, last interval to 1e10."""
    g = golden(f'tensorf_world_{mode}')
    configs, model_configs = golden_configs('tensorf_world')
    assert configs['data_loader']['ndc'] is False
    sets = FX.tensorf_sets(configs, seed=int(g['param_seed']), with_alpha=bool(g['with_alpha']))
    torch.manual_seed(int(g['rng_seed']))
    with torch.no_grad():
        out = P.tensorf_render_chunk(sets, configs, model_configs, g['pixel_id'], training=(mode == 'train'))
    for k in ('rays_o', 'rays_d', 'view_dirs', 'z_vals_coarse'):
        assert torch.equal(out[k], g[k]), k
    assert torch.equal(out['validity_mask_coarse'], g['validity_mask_coarse'])
    assert torch.equal(out['surface_mask_coarse'], g['surface_mask_coarse'])
    assert not any('ndc' in k for k in out)
    for k in g:
        if k in out and g[k].dtype == torch.float32:
            assert _close(out[k], g[k], 2e-4 * max(1.0, g[k].abs().max().item())), k


@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_tensorf_variant_pipeline_matches_reference_golden(golden, golden_configs, mode):
    """SoftPlus density, another distance scale, white background, a view-independent colour predictor on the augmentation tensor."""
    g = golden(f'tensorf_variant_{mode}')
    configs, model_configs = golden_configs('tensorf_variant')
    assert configs['model']['coarse_model']['density_predictor'] == 'SoftPlus' and configs['model']['white_bkgd']
    assert configs['model']['augmentations'][0]['coarse_model']['use_view_dirs'] is False
    sets = FX.tensorf_sets(configs, seed=int(g['param_seed']), with_alpha=bool(g['with_alpha']))
    assert sets['augmentations'][0][2]['params']['color_predictor.mlp.0.weight'].shape[1] == 27
    torch.manual_seed(int(g['rng_seed']))
    with torch.no_grad():
        out = P.tensorf_render_chunk(sets, configs, model_configs, g['pixel_id'], training=(mode == 'train'))
    assert torch.equal(out['z_vals_coarse'], g['z_vals_coarse'])
    assert torch.equal(out['validity_mask_coarse'], g['validity_mask_coarse'])
    for k in g:
        if k in out and g[k].dtype == torch.float32:
            assert _close(out[k], g[k], 2e-4 * max(1.0, g[k].abs().max().item())), k


@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_tensorf_cp_pipeline_matches_reference_golden(golden, golden_configs, mode):
    """`decomposition_type = "CandecompParafac"` (SimpleTensoRF09.py:964-1124): three line factors per component."""
    g = golden(f'tensorf_cp_{mode}')
    configs, model_configs = golden_configs('tensorf_cp')
    assert configs['model']['coarse_model']['decomposition_type'] == 'CandecompParafac'
    sets = FX.tensorf_sets(configs, seed=int(g['param_seed']), with_alpha=bool(g['with_alpha']))
    assert TF.is_cp(sets['coarse_model']['params'])
    torch.manual_seed(int(g['rng_seed']))
    with torch.no_grad():
        out = P.tensorf_render_chunk(sets, configs, model_configs, g['pixel_id'], training=(mode == 'train'))
    assert torch.equal(out['z_vals_coarse'], g['z_vals_coarse'])
    assert torch.equal(out['validity_mask_coarse'], g['validity_mask_coarse'])
    assert torch.equal(out['surface_mask_coarse'], g['surface_mask_coarse'])
    assert 0.2 < g['surface_mask_coarse'].float().mean() < 0.8                       # the fixture exercises the colour branch
    for k in g:
        if k in out and g[k].dtype == torch.float32:
            assert _close(out[k], g[k], 2e-4 * max(1.0, g[k].abs().max().item())), k


@pytest.mark.parametrize('n', [1, 3, 5, 7, 8, 35, 62, 64, 190, 462, 1081, 2312])
def test_aten_row_sum_order(n):
    torch.manual_seed(n)
    x = torch.rand(40, n) ** 3 + 1e-5
    ref = x.sum(-1).numpy()
    for r in range(x.shape[0]):
        assert AO.row_sum_f32(x[r].numpy()) == ref[r]


def test_aten_cumsum_order():
    torch.manual_seed(1)
    x = torch.rand(50, 62) ** 6
    x = x / x.sum(-1, keepdim=True)
    ref = torch.cumsum(x, -1).numpy()
    for r in range(x.shape[0]):
        assert np.array_equal(AO.cumsum_f32(x[r].numpy()), ref[r])


def test_inverse_cdf_row_restatement_is_bit_exact(golden):
    g = golden('sample_pdf')
    for tag in 'abc':
        z, w, u = g[f'{tag}_z'], g[f'{tag}_weights'], g[f'{tag}_u']
        mids = (.5 * (z[..., 1:] + z[..., :-1])).numpy()
        for r in range(0, z.shape[0], 5):
            s, b, a, _ = AO.inverse_cdf_row(mids[r], w[r, 1:-1].numpy(), u[r].numpy())
            assert np.array_equal(s, g[f'{tag}_samples'][r].numpy())
            assert np.array_equal(b, g[f'{tag}_below'][r].numpy()) and np.array_equal(a, g[f'{tag}_above'][r].numpy())


def test_trilinear_positive_restatement():
    g = torch.Generator().manual_seed(3)
    Z, Y, X = 7, 9, 11
    vol = (torch.rand(Z, Y, X, generator=g) < 0.3).float()
    bbox = torch.tensor([[-1.5, -1.67, -1.0], [1.5, 1.67, 1.0]])
    pts = (torch.rand(600, 3, generator=g) * 1.1 - 0.05) * (bbox[1] - bbox[0]) + bbox[0]
    # points exactly on voxel planes and on the box faces
    lat = torch.stack([torch.linspace(-1.5, 1.5, X)[torch.randint(0, X, (100,), generator=g)],
                       torch.linspace(-1.67, 1.67, Y)[torch.randint(0, Y, (100,), generator=g)],
                       torch.linspace(-1.0, 1.0, Z)[torch.randint(0, Z, (100,), generator=g)]], 1)
    pts = torch.cat([pts, lat, bbox[0][None], bbox[1][None]])
    ref = TF.sample_alpha(vol.view(1, 1, Z, Y, X), bbox, pts) > 0
    mine = AO.trilinear_positive(vol.numpy(), bbox.numpy(), pts.numpy())
    assert np.array_equal(mine, ref.numpy())


def _probe_loss(out, keys, seed):
    g = torch.Generator().manual_seed(seed)
    return sum((out[k] * torch.randn(out[k].shape, generator=g)).sum() for k in keys)


def test_learnable_camera_gradients_match_reference_golden(golden, golden_configs):
    """Pose-correction gradients (SimpleNeRF17.py:817-842) of a fixed linear functional of the rendered maps: the oracle pipeline fed with
    `rays.pose_correction(...)` against r.grad / t.grad of the unmodified reference models (oracle/generate_golden.py::golden_learnable_cameras)."""
    from oracle import rays as RY
    g = golden('learnable_cameras')
    # Simple-NeRF: NDC, coarse + fine + augmentation MLPs in training mode
    configs, model_configs = golden_configs('nerf')
    sets = FX.nerf_param_sets(configs, seed=11)
    r, t = g['nerf_r'].clone().requires_grad_(), g['nerf_t'].clone().requires_grad_()
    torch.manual_seed(909)
    out = P.nerf_render_chunk(sets, configs, model_configs, g['nerf_pixel_id'], training=True,
                              extrinsics=RY.pose_correction(torch.tensor(model_configs['extrinsics']), r, t))
    keys = ('rgb_coarse', 'rgb_fine', 'depth_coarse', 'depth_fine', 'depth_ndc_fine', 'acc_fine', 'rays_o', 'rays_d_ndc', 'view_dirs')
    _probe_loss(out, keys, 77).backward()
    assert _close(out['rgb_fine'].detach(), g['nerf_rgb_fine'], 2e-4)
    for got, want in ((r.grad, g['nerf_r_grad']), (t.grad, g['nerf_t_grad'])):
        assert float(want.abs().max()) > 0
        assert _close(got, want, 2e-3 * float(want.abs().max())), float((got - want).abs().max() / want.abs().max())
    # Simple-TensoRF: NDC; the grid coordinates are detached upstream, the pose is reached through view_dirs, |d| and the world depths
    configs, model_configs = golden_configs('tensorf')
    sets = FX.tensorf_sets(configs, seed=21, with_alpha=False)
    r, t = g['tensorf_r'].clone().requires_grad_(), g['tensorf_t'].clone().requires_grad_()
    torch.manual_seed(910)
    out = P.tensorf_render_chunk(sets, configs, model_configs, g['tensorf_pixel_id'], training=True,
                                 extrinsics=RY.pose_correction(torch.tensor(model_configs['extrinsics']), r, t))
    _probe_loss(out, ('rgb_coarse', 'depth_coarse', 'depth_ndc_coarse', 'acc_coarse', 'view_dirs', 'rays_d'), 78).backward()
    assert _close(out['rgb_coarse'].detach(), g['tensorf_rgb'], 2e-4)
    for got, want in ((r.grad, g['tensorf_r_grad']), (t.grad, g['tensorf_t_grad'])):
        assert float(want.abs().max()) > 0
        assert _close(got, want, 2e-3 * float(want.abs().max())), float((got - want).abs().max() / want.abs().max())


def test_learnable_camera_gradients_world_space_nerf(golden, golden_configs):
    """Simple-NeRF with `ndc = False`, lindisp depths, white background and learnable cameras (oracle/generate_golden.py::nerf_variant_configs)."""
    from oracle import rays as RY
    g = golden('learnable_cameras')
    configs, model_configs = golden_configs('nerf_variant')
    sets = FX.nerf_param_sets(configs, seed=13)
    r, t = g['nerf_world_r'].clone().requires_grad_(), g['nerf_world_t'].clone().requires_grad_()
    torch.manual_seed(912)
    out = P.nerf_render_chunk(sets, configs, model_configs, g['nerf_world_pixel_id'], training=True,
                              extrinsics=RY.pose_correction(torch.tensor(model_configs['extrinsics']), r, t))
    _probe_loss(out, ('rgb_coarse', 'rgb_fine', 'depth_coarse', 'depth_fine', 'depth_var_fine', 'acc_fine', 'rays_d', 'view_dirs'), 80).backward()
    assert _close(out['rgb_fine'].detach(), g['nerf_world_rgb_fine'], 2e-4)
    for got, want in ((r.grad, g['nerf_world_r_grad']), (t.grad, g['nerf_world_t_grad'])):
        assert float(want.abs().max()) > 0
        assert _close(got, want, 2e-3 * float(want.abs().max())), float((got - want).abs().max() / want.abs().max())


def test_learnable_camera_gradients_world_space_tensorf(golden, golden_configs):
    """Simple-TensoRF without NDC: the box-march depths start where the ray enters the box (SimpleTensoRF09.py:388-400), so the pose is also
    reached through z (depth, depth_var, the intervals)."""
    from oracle import rays as RY
    g = golden('learnable_cameras')
    configs, model_configs = golden_configs('tensorf_world')
    sets = FX.tensorf_sets(configs, seed=23, with_alpha=False)
    r, t = g['tensorf_world_r'].clone().requires_grad_(), g['tensorf_world_t'].clone().requires_grad_()
    torch.manual_seed(911)
    out = P.tensorf_render_chunk(sets, configs, model_configs, g['tensorf_world_pixel_id'], training=True,
                                 extrinsics=RY.pose_correction(torch.tensor(model_configs['extrinsics']), r, t))
    _probe_loss(out, ('rgb_coarse', 'depth_coarse', 'depth_var_coarse', 'acc_coarse', 'view_dirs', 'rays_d', 'z_vals_coarse'), 79).backward()
    assert _close(out['rgb_coarse'].detach(), g['tensorf_world_rgb'], 2e-4)
    for got, want in ((r.grad, g['tensorf_world_r_grad']), (t.grad, g['tensorf_world_t_grad'])):
        assert float(want.abs().max()) > 0
        assert _close(got, want, 2e-3 * float(want.abs().max())), float((got - want).abs().max() / want.abs().max())


def test_pose_correction_matches_the_dropin_learner(golden_configs):
    """oracle.rays.pose_correction == the drop-in's ExtrinsicsLearner.forward (the class Trainer10 / Tester07 optimise), values and gradients."""
    from oracle import rays as RY
    from simple_rf_b200.models.SimpleNeRF91 import ExtrinsicsLearner
    _, model_configs = golden_configs('nerf')
    E = torch.tensor(model_configs['extrinsics']).float()
    learner = ExtrinsicsLearner(E.numpy(), learn_rotation=True, learn_translation=True)
    gen = torch.Generator().manual_seed(3)
    learner.r.data.copy_(torch.randn(learner.r.shape, generator=gen) * 0.05)
    learner.t.data.copy_(torch.randn(learner.t.shape, generator=gen) * 0.05)
    r, t = learner.r.detach().clone().requires_grad_(), learner.t.detach().clone().requires_grad_()
    a = learner(torch.arange(learner.num_frames))
    b = RY.pose_correction(E, r, t)
    assert torch.allclose(a, b, rtol=0, atol=1e-6)
    probe = torch.randn(a.shape, generator=gen)
    (a * probe).sum().backward()
    (b * probe).sum().backward()
    assert torch.allclose(learner.r.grad, r.grad, rtol=1e-5, atol=1e-6) and torch.allclose(learner.t.grad, t.grad, rtol=1e-5, atol=1e-6)
