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
