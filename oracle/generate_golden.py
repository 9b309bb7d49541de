"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only) and pin
the oracle against it on the way.   python -m oracle.generate_golden

Every fixture stores the seeded inputs and the REFERENCE's outputs; model parameters are not
stored but regenerated from a seeded torch.Generator by oracle.nerf_mlp.init_mlp_params /
oracle.tensorf.init_vm_params on both sides (they are loaded into the reference modules here).
The script fails if any oracle stage differs from the reference (bit-exact on this host, because
both run the same ATen CPU kernels in the same order).
"""
import copy
import json
import sys
from pathlib import Path

import numpy as np
import torch

from . import composite as C
from . import losses as OL
from . import fixtures as FX
from . import nerf_mlp as M
from . import pipeline as P
from . import reference_harness as H
from . import sampling as SP
from . import surgery as SG
from . import tensorf as TF

OUT = Path(__file__).resolve().parent.parent / 'tests' / 'golden'


def _np(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def _check(name, ref, mine, exact=True, tol=0.0):
    ref, mine = torch.as_tensor(ref), torch.as_tensor(mine)
    assert ref.shape == mine.shape, (name, ref.shape, mine.shape)
    if exact:
        ok = torch.equal(ref, mine)
    else:
        ok = bool((ref.double() - mine.double()).abs().max() <= tol)
    if not ok:
        raise SystemExit(f'oracle != reference at {name}: max abs diff '
                         f'{(ref.double() - mine.double()).abs().max().item():.3e}')


def load_nerf_params(model, sets):
    def put(module, params):
        sd = module.state_dict()
        assert set(sd.keys()) == set(params.keys()), (sorted(sd.keys()), sorted(params.keys()))
        module.load_state_dict(params)
    put(model.coarse_model, sets['coarse_model'])
    put(model.fine_model, sets['fine_model'])
    for aug, (_, _, params) in zip(model.augmented_models, sets['augmentations']):
        put(aug['coarse_model'], params)


def golden_nerf():
    configs, model_configs = H.load_configs(1142, 'fern')
    model_configs = H.shrink(model_configs, 4)
    configs['model']['netchunk'] = 2048
    (OUT / 'nerf_configs.json').write_text(json.dumps({'configs': configs, 'model_configs': model_configs}, indent=1))
    model = H.build_model(configs, model_configs)
    sets = FX.nerf_param_sets(configs, seed=11)
    load_nerf_params(model, sets)
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    keep = ('rays_o', 'rays_d', 'rays_o_ndc', 'rays_d_ndc', 'view_dirs')
    per_tag = ('rgb', 'acc', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc', 'alpha', 'visibility',
               'weights', 'raw_sigma', 'raw_rgb')
    for mode, R, seed in (('eval', 48, 3), ('train', 40, 4)):
        pixel_id = FX.random_pixels(R, nviews, h, w, seed)
        model.train(mode == 'train')
        torch.manual_seed(100 + seed)
        with torch.no_grad():
            ref = model({'pixel_id': pixel_id, 'num_frames': nviews, 'iter_num': 0, 'sub_batch_index': 0}, retraw=True)
        torch.manual_seed(100 + seed)
        with torch.no_grad():
            mine = P.nerf_render_chunk(sets, configs, model_configs, pixel_id, training=(mode == 'train'))
        fixture = {'pixel_id': pixel_id, 'param_seed': 11, 'rng_seed': 100 + seed}
        for k in keep + ('z_vals_coarse', 'z_vals_fine'):
            _check(f'nerf/{mode}/{k}', ref[k], mine[k])
            fixture[k] = ref[k]
        prefixes = [''] + ([f"{a[0]}_" for a in sets['augmentations']] if mode == 'train' else [])
        for pre in prefixes:
            for tag in ('coarse', 'fine'):
                if pre and tag == 'fine':
                    continue
                for k in per_tag:
                    key = f'{pre}{k}_{tag}'
                    _check(f'nerf/{mode}/{key}', ref[key], mine[key])
                    fixture[key] = ref[key]
        for k in ('_fine_u', '_fine_below', '_fine_above', '_fine_samples'):
            fixture[k[1:]] = mine[k]
        np.savez_compressed(OUT / f'nerf_{mode}.npz', **_np(fixture))
        print(f'nerf_{mode}: {len(fixture)} arrays, oracle == reference')
    return configs, model_configs


def nerf_variant_configs():
    """The shipped train1142 config with the switches no shipped run flips: world-space sampling (`data_loader.ndc = False`, near / far in
    world units), depths linear in disparity (`lindisp`), white background, a smaller noise level."""
    configs, model_configs = H.load_configs(1142, 'fern')
    model_configs = H.shrink(model_configs, 4)
    configs['model']['netchunk'] = 2048
    configs['data_loader']['ndc'] = False
    configs['model']['lindisp'] = True
    configs['model']['white_bkgd'] = True
    configs['model']['raw_noise_std'] = 0.25
    return configs, model_configs


def golden_nerf_variants():
    """Simple-NeRF in world space with lindisp depths and a white background through the unmodified reference, eval and train."""
    configs, model_configs = nerf_variant_configs()
    (OUT / 'nerf_variant_configs.json').write_text(json.dumps({'configs': configs, 'model_configs': model_configs}, indent=1))
    model = H.build_model(configs, model_configs)
    sets = FX.nerf_param_sets(configs, seed=13)
    load_nerf_params(model, sets)
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    per_tag = ('rgb', 'acc', 'depth', 'depth_var', 'alpha', 'visibility', 'weights', 'raw_sigma', 'raw_rgb')
    for mode, R, seed in (('eval', 48, 7), ('train', 40, 8)):
        pixel_id = FX.random_pixels(R, nviews, h, w, seed)
        model.train(mode == 'train')
        torch.manual_seed(600 + seed)
        with torch.no_grad():
            ref = model({'pixel_id': pixel_id, 'num_frames': nviews, 'iter_num': 0, 'sub_batch_index': 0}, retraw=True)
        torch.manual_seed(600 + seed)
        with torch.no_grad():
            mine = P.nerf_render_chunk(sets, configs, model_configs, pixel_id, training=(mode == 'train'))
        assert 'rays_o_ndc' not in ref and 'depth_ndc_coarse' not in ref
        fixture = {'pixel_id': pixel_id, 'param_seed': 13, 'rng_seed': 600 + seed}
        for k in ('rays_o', 'rays_d', 'view_dirs', 'z_vals_coarse', 'z_vals_fine'):
            _check(f'nerf_variant/{mode}/{k}', ref[k], mine[k])
            fixture[k] = ref[k]
        prefixes = [''] + ([f"{a[0]}_" for a in sets['augmentations']] if mode == 'train' else [])
        for pre in prefixes:
            for tag in ('coarse', 'fine'):
                if pre and tag == 'fine':
                    continue
                for k in per_tag:
                    key = f'{pre}{k}_{tag}'
                    _check(f'nerf_variant/{mode}/{key}', ref[key], mine[key])
                    fixture[key] = ref[key]
        np.savez_compressed(OUT / f'nerf_variant_{mode}.npz', **_np(fixture))
        print(f'nerf_variant_{mode}: {len(fixture)} arrays, oracle == reference (acc mean {ref["acc_fine"].mean():.3f})')


def golden_sample_pdf():
    """Stage-wise: the reference's own static sample_pdf + get_z_vals_fine on seeded inputs, both
    the deterministic (eval) and the random-u (train) form, S in {64, 37}."""
    get_model, _ = H.import_reference()
    from models.SimpleNeRF17 import SimpleNeRF
    fixture = {}
    for tag, S, N, R, det in (('a', 64, 128, 64, True), ('b', 64, 128, 64, False), ('c', 37, 50, 33, False)):
        g = torch.Generator().manual_seed(7 + S + N + det)
        z = torch.sort(torch.rand(R, S, generator=g), -1)[0]
        wts = torch.rand(R, S, generator=g) ** 4
        wts[: R // 4] *= 1e-4                      # nearly-empty rays: cdf dominated by the 1e-5 floor
        wts[R // 4: R // 2, S // 3:] = 0           # mass concentrated early: many repeated cdf values
        if det:
            torch.manual_seed(0)
            u = SP.det_u(R, N)
        else:
            torch.manual_seed(31 + S)
            u = torch.rand([R, N])
            torch.manual_seed(31 + S)
        mids = .5 * (z[..., 1:] + z[..., :-1])
        ref_samples = SimpleNeRF.sample_pdf(mids, wts[..., 1:-1], N, det=det)
        ref_fine = torch.sort(torch.cat([z, ref_samples], -1), -1)[0]
        z_f, samples, below, above = SP.fine_depths(z, wts, u)
        _check(f'sample_pdf/{tag}/samples', ref_samples, samples)
        _check(f'sample_pdf/{tag}/z_fine', ref_fine, z_f)
        fixture.update({f'{tag}_z': z, f'{tag}_weights': wts, f'{tag}_u': u, f'{tag}_samples': ref_samples,
                        f'{tag}_z_fine': ref_fine, f'{tag}_below': below, f'{tag}_above': above})
    np.savez_compressed(OUT / 'sample_pdf.npz', **_np(fixture))
    print('sample_pdf: oracle == reference')


def golden_composite(configs, model_configs):
    """Stage-wise: reference volume_rendering (NDC and world) with autograd gradients."""
    model = H.build_model(configs, model_configs)
    fixture = {}
    for tag, ndc, white, S in (('ndc', True, False, 64), ('world', False, True, 45)):
        g = torch.Generator().manual_seed(5 + S)
        R = 24
        model.ndc = ndc
        model.configs['model']['white_bkgd'] = white
        sigma = (torch.relu(torch.randn(R, S, generator=g)) * 10).requires_grad_()
        rgb = torch.rand(R, S, 3, generator=g).requires_grad_()
        z = torch.sort(torch.rand(R, S, generator=g), -1)[0]
        if not ndc:
            z = 2 + 4 * z
        rays_o = torch.randn(R, 3, generator=g) * 0.1
        rays_d = torch.randn(R, 3, generator=g) * 0.3 - torch.tensor([0, 0, 1.])
        d_ndc = torch.randn(R, 3, generator=g)
        net = {'sigma': sigma[..., None], 'rgb': rgb}
        if ndc:
            ref = model.volume_rendering(net, z_vals_ndc=z, rays_d_ndc=d_ndc, rays_o=rays_o, rays_d=rays_d)
        else:
            ref = model.volume_rendering(net, z_vals=z, rays_d=rays_d)
        mine = C.composite(sigma, rgb, z, rays_o, rays_d, d_ndc, ndc=ndc, white_bkgd=white)
        ups = {}
        loss = 0
        for k in ('rgb', 'acc', 'depth', 'depth_var', 'weights') + (('depth_ndc', 'depth_var_ndc') if ndc else ()):
            _check(f'composite/{tag}/{k}', ref[k], mine[k])
            ups[k] = torch.rand(ref[k].shape, generator=g)
            loss = loss + (ref[k] * ups[k]).sum()
        gs, gc = torch.autograd.grad(loss, [sigma, rgb])
        gs_o, gc_o = C.composite_backward(
            sigma.detach().double(), rgb.detach().double(), z.double(), rays_o.double(), rays_d.double(),
            d_ndc.double(), ndc=ndc, white_bkgd=white, g_rgb=ups['rgb'].double(), g_acc=ups['acc'].double(),
            g_depth=ups['depth'].double(), g_depth_var=ups['depth_var'].double(),
            g_depth_ndc=ups['depth_ndc'].double() if ndc else None,
            g_depth_var_ndc=ups['depth_var_ndc'].double() if ndc else None, g_weights=ups['weights'].double())
        scale = gs.abs().max().item()
        _check(f'composite/{tag}/g_sigma', gs / scale, gs_o.float() / scale, exact=False, tol=2e-4)
        _check(f'composite/{tag}/g_rgb', gc, gc_o.float(), exact=False, tol=1e-5)
        fixture.update({f'{tag}_sigma': sigma, f'{tag}_rgb': rgb, f'{tag}_z': z, f'{tag}_rays_o': rays_o,
                        f'{tag}_rays_d': rays_d, f'{tag}_rays_d_ndc': d_ndc, f'{tag}_g_sigma': gs, f'{tag}_g_rgb': gc})
        for k, v in ups.items():
            fixture[f'{tag}_up_{k}'] = v
        for k in ('rgb', 'acc', 'depth', 'depth_var', 'weights', 'alpha', 'visibility') + (('depth_ndc', 'depth_var_ndc') if ndc else ()):
            fixture[f'{tag}_out_{k}'] = ref[k]
    np.savez_compressed(OUT / 'composite.npz', **_np(fixture))
    print('composite: oracle == reference (forward exact, closed-form backward vs autograd)')


def load_tensorf_params(model, sets):
    from models.SimpleTensoRF09 import AlphaGridMask

    def put(module, t):
        sd = dict(module.named_parameters())
        assert set(sd.keys()) == set(t['params'].keys()), (sorted(sd.keys()), sorted(t['params'].keys()))
        for k, v in t['params'].items():
            assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
            sd[k].data.copy_(v)
        assert int(module.num_samples) == t['num_samples']
        module.alpha_mask = AlphaGridMask(t['alpha_volume'][0, 0], t['alpha_bbox']) if 'alpha_volume' in t else None
    put(model.coarse_model, sets['coarse_model'])
    for aug, (_, _, t) in zip(model.augmented_models, sets['augmentations']):
        put(aug['coarse_model'], t)


def golden_tensorf():
    configs, model_configs = H.load_configs(212, '00000')
    model_configs = H.shrink(model_configs, 4)
    configs['model']['coarse_model']['num_voxels_initial'] = 40 ** 3
    configs['model']['augmentations'][0]['coarse_model']['num_voxels_initial'] = 20 ** 3
    (OUT / 'tensorf_configs.json').write_text(json.dumps({'configs': configs, 'model_configs': model_configs}, indent=1))
    model = H.build_model(configs, model_configs)
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    keys = ('rgb', 'acc', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc', 'alpha', 'visibility', 'weights',
            'raw_sigma', 'raw_rgb')
    for mode, R, seed, with_alpha in (('eval', 40, 5, True), ('train', 32, 6, False)):
        sets = FX.tensorf_sets(configs, seed=21, with_alpha=with_alpha)
        load_tensorf_params(model, sets)
        pixel_id = FX.random_pixels(R, nviews, h, w, seed)
        model.train(mode == 'train')
        torch.manual_seed(200 + seed)
        with torch.no_grad():
            ref = model({'pixel_id': pixel_id, 'num_frames': nviews, 'iter_num': 1, 'sub_batch_index': 1}, retraw=True)
        torch.manual_seed(200 + seed)
        with torch.no_grad():
            mine = P.tensorf_render_chunk(sets, configs, model_configs, pixel_id, training=(mode == 'train'))
        fixture = {'pixel_id': pixel_id, 'param_seed': 21, 'rng_seed': 200 + seed, 'with_alpha': with_alpha}
        for k in ('rays_o', 'rays_d', 'rays_o_ndc', 'rays_d_ndc', 'view_dirs', 'z_vals_coarse'):
            _check(f'tensorf/{mode}/{k}', ref[k], mine[k])
            fixture[k] = ref[k]
        prefixes = [''] + ([f"{a[0]}_" for a in sets['augmentations']] if mode == 'train' else [])
        for pre in prefixes:
            for k in keys:
                key = f'{pre}{k}_coarse'
                _check(f'tensorf/{mode}/{key}', ref[key], mine[key])
                fixture[key] = ref[key]
            fixture[f'{pre}validity_mask_coarse'] = mine[f'{pre}validity_mask_coarse']
            fixture[f'{pre}surface_mask_coarse'] = mine[f'{pre}surface_mask_coarse']
        frac = mine['validity_mask_coarse'].float().mean().item(), mine['surface_mask_coarse'].float().mean().item()
        np.savez_compressed(OUT / f'tensorf_{mode}.npz', **_np(fixture))
        print(f'tensorf_{mode}: oracle == reference (valid {frac[0]:.3f}, surface {frac[1]:.3f})')


def as_cp(configs, comps_density=24, comps_color=48):
    """Switch every tensor of a Simple-TensoRF config to the CANDECOMP/PARAFAC decomposition (SimpleTensoRF09.py:537-539; selected
    by no shipped config).  One component count per tensor: every line holds num_components[0] components (:992) and
    basis_matrix_color takes sum(num_components_color) inputs (:986)."""
    tensors = [configs['model']['coarse_model']] + [a['coarse_model'] for a in configs['model'].get('augmentations', [])]
    for cfg in tensors:
        cfg['decomposition_type'] = 'CandecompParafac'
        cfg['num_components_density'] = [comps_density]
        cfg['num_components_color'] = [comps_color]
    return configs


def golden_tensorf_cp():
    """The CP tensor through the unmodified reference: eval (with an alpha mask) and train (jitter, augmentation tensor)."""
    configs, model_configs = H.load_configs(212, '00000')
    model_configs = H.shrink(model_configs, 4)
    as_cp(configs)
    configs['model']['coarse_model']['num_voxels_initial'] = 40 ** 3
    configs['model']['augmentations'][0]['coarse_model']['num_voxels_initial'] = 20 ** 3
    (OUT / 'tensorf_cp_configs.json').write_text(json.dumps({'configs': configs, 'model_configs': model_configs}, indent=1))
    model = H.build_model(configs, model_configs)
    assert type(model.coarse_model).__name__ == 'CpDecomposedTensor'
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    keys = ('rgb', 'acc', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc', 'alpha', 'visibility', 'weights',
            'raw_sigma', 'raw_rgb')
    for mode, R, seed, with_alpha in (('eval', 40, 25, True), ('train', 32, 26, False)):
        sets = FX.tensorf_sets(configs, seed=27, with_alpha=with_alpha)
        load_tensorf_params(model, sets)
        pixel_id = FX.random_pixels(R, nviews, h, w, seed)
        model.train(mode == 'train')
        torch.manual_seed(500 + seed)
        with torch.no_grad():
            ref = model({'pixel_id': pixel_id, 'num_frames': nviews, 'iter_num': 1, 'sub_batch_index': 1}, retraw=True)
        torch.manual_seed(500 + seed)
        with torch.no_grad():
            mine = P.tensorf_render_chunk(sets, configs, model_configs, pixel_id, training=(mode == 'train'))
        fixture = {'pixel_id': pixel_id, 'param_seed': 27, 'rng_seed': 500 + seed, 'with_alpha': with_alpha}
        for k in ('rays_o', 'rays_d', 'rays_o_ndc', 'rays_d_ndc', 'view_dirs', 'z_vals_coarse'):
            _check(f'tensorf_cp/{mode}/{k}', ref[k], mine[k])
            fixture[k] = ref[k]
        prefixes = [''] + ([f"{a[0]}_" for a in sets['augmentations']] if mode == 'train' else [])
        for pre in prefixes:
            for k in keys:
                key = f'{pre}{k}_coarse'
                _check(f'tensorf_cp/{mode}/{key}', ref[key], mine[key])
                fixture[key] = ref[key]
            fixture[f'{pre}validity_mask_coarse'] = mine[f'{pre}validity_mask_coarse']
            fixture[f'{pre}surface_mask_coarse'] = mine[f'{pre}surface_mask_coarse']
        frac = mine['validity_mask_coarse'].float().mean().item(), mine['surface_mask_coarse'].float().mean().item()
        np.savez_compressed(OUT / f'tensorf_cp_{mode}.npz', **_np(fixture))
        print(f'tensorf_cp_{mode}: oracle == reference (valid {frac[0]:.3f}, surface {frac[1]:.3f}, acc mean {ref["acc_coarse"].mean():.3f})')


def tensorf_variant_configs():
    """The shipped train0212 config with the switches no shipped run flips: SoftPlus density (`density_offset` -1), another `distance_scale`,
    a white background, and a view-INdependent colour predictor on the augmentation tensor (`use_view_dirs` / `view_dependent_color` false,
    SimpleTensoRF09.py:732, :1267, :1384, :1414)."""
    configs, model_configs = H.load_configs(212, '00000')
    model_configs = H.shrink(model_configs, 4)
    main, aug = configs['model']['coarse_model'], configs['model']['augmentations'][0]['coarse_model']
    main['num_voxels_initial'], aug['num_voxels_initial'] = 40 ** 3, 20 ** 3
    main.update(density_predictor='SoftPlus', density_offset=-1.0, distance_scale=10)
    aug.update(use_view_dirs=False, view_dependent_color=False)
    configs['model']['white_bkgd'] = True
    return configs, model_configs


def golden_tensorf_variants():
    configs, model_configs = tensorf_variant_configs()
    (OUT / 'tensorf_variant_configs.json').write_text(json.dumps({'configs': configs, 'model_configs': model_configs}, indent=1))
    model = H.build_model(configs, model_configs)
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    keys = ('rgb', 'acc', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc', 'alpha', 'visibility', 'weights', 'raw_sigma', 'raw_rgb')
    for mode, R, seed, with_alpha in (('eval', 40, 35, True), ('train', 32, 36, False)):
        sets = FX.tensorf_sets(configs, seed=29, with_alpha=with_alpha)
        load_tensorf_params(model, sets)
        pixel_id = FX.random_pixels(R, nviews, h, w, seed)
        model.train(mode == 'train')
        torch.manual_seed(700 + seed)
        with torch.no_grad():
            ref = model({'pixel_id': pixel_id, 'num_frames': nviews, 'iter_num': 1, 'sub_batch_index': 1}, retraw=True)
        torch.manual_seed(700 + seed)
        with torch.no_grad():
            mine = P.tensorf_render_chunk(sets, configs, model_configs, pixel_id, training=(mode == 'train'))
        fixture = {'pixel_id': pixel_id, 'param_seed': 29, 'rng_seed': 700 + seed, 'with_alpha': with_alpha}
        for k in ('rays_o', 'rays_d', 'rays_o_ndc', 'rays_d_ndc', 'view_dirs', 'z_vals_coarse'):
            _check(f'tensorf_variant/{mode}/{k}', ref[k], mine[k])
            fixture[k] = ref[k]
        prefixes = [''] + ([f"{a[0]}_" for a in sets['augmentations']] if mode == 'train' else [])
        for pre in prefixes:
            for k in keys:
                key = f'{pre}{k}_coarse'
                _check(f'tensorf_variant/{mode}/{key}', ref[key], mine[key])
                fixture[key] = ref[key]
            fixture[f'{pre}validity_mask_coarse'] = mine[f'{pre}validity_mask_coarse']
            fixture[f'{pre}surface_mask_coarse'] = mine[f'{pre}surface_mask_coarse']
        frac = mine['surface_mask_coarse'].float().mean().item()
        np.savez_compressed(OUT / f'tensorf_variant_{mode}.npz', **_np(fixture))
        print(f'tensorf_variant_{mode}: oracle == reference (surface {frac:.3f}, acc mean {ref["acc_coarse"].mean():.3f})')


def tensorf_world_configs():
    """The shipped train0212 config with `ndc = False` (no shipped run selects it; SimpleTensoRF09.py:388-400 is the box-marching
    sampler it switches on): world-space tensor boxes in front of the cameras, near / far in world units."""
    configs, model_configs = H.load_configs(212, '00000')
    model_configs = H.shrink(model_configs, 4)
    configs['data_loader']['ndc'] = False
    box = [[-2.0, -1.8, -7.0], [2.0, 1.8, -1.0]]
    configs['model']['coarse_model']['num_voxels_initial'] = 40 ** 3
    configs['model']['coarse_model']['bounding_box'] = box
    configs['model']['augmentations'][0]['coarse_model']['num_voxels_initial'] = 20 ** 3
    configs['model']['augmentations'][0]['coarse_model']['bounding_box'] = box
    model_configs['near'], model_configs['far'] = 0.75, 9.0
    return configs, model_configs


def golden_tensorf_world():
    """Simple-TensoRF without NDC: box-march depths (bit-exact), world-space points / view directions, compositing with the last
    interval to 1e10 — the unmodified reference, eval (with an alpha mask) and train (per-ray jitter, augmentation tensor)."""
    configs, model_configs = tensorf_world_configs()
    (OUT / 'tensorf_world_configs.json').write_text(json.dumps({'configs': configs, 'model_configs': model_configs}, indent=1))
    model = H.build_model(configs, model_configs)
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    keys = ('rgb', 'acc', 'depth', 'depth_var', 'alpha', 'visibility', 'weights', 'raw_sigma', 'raw_rgb')
    for mode, R, seed, with_alpha in (('eval', 40, 15, True), ('train', 32, 16, False)):
        sets = FX.tensorf_sets(configs, seed=23, with_alpha=with_alpha)
        load_tensorf_params(model, sets)
        pixel_id = FX.random_pixels(R, nviews, h, w, seed)
        model.train(mode == 'train')
        torch.manual_seed(400 + seed)
        with torch.no_grad():
            ref = model({'pixel_id': pixel_id, 'num_frames': nviews, 'iter_num': 1, 'sub_batch_index': 1}, retraw=True)
        torch.manual_seed(400 + seed)
        with torch.no_grad():
            mine = P.tensorf_render_chunk(sets, configs, model_configs, pixel_id, training=(mode == 'train'))
        assert 'depth_ndc_coarse' not in ref and 'rays_o_ndc' not in ref
        fixture = {'pixel_id': pixel_id, 'param_seed': 23, 'rng_seed': 400 + seed, 'with_alpha': with_alpha}
        for k in ('rays_o', 'rays_d', 'view_dirs', 'z_vals_coarse'):
            _check(f'tensorf_world/{mode}/{k}', ref[k], mine[k])
            fixture[k] = ref[k]
        prefixes = [''] + ([f"{a[0]}_" for a in sets['augmentations']] if mode == 'train' else [])
        for pre in prefixes:
            for k in keys:
                key = f'{pre}{k}_coarse'
                _check(f'tensorf_world/{mode}/{key}', ref[key], mine[key])
                fixture[key] = ref[key]
            fixture[f'{pre}validity_mask_coarse'] = mine[f'{pre}validity_mask_coarse']
            fixture[f'{pre}surface_mask_coarse'] = mine[f'{pre}surface_mask_coarse']
        frac = mine['validity_mask_coarse'].float().mean().item(), mine['surface_mask_coarse'].float().mean().item()
        np.savez_compressed(OUT / f'tensorf_world_{mode}.npz', **_np(fixture))
        print(f'tensorf_world_{mode}: oracle == reference (valid {frac[0]:.3f}, surface {frac[1]:.3f}, acc mean {ref["acc_coarse"].mean():.3f})')


def tensorf_full_size_configs():
    """Shipped train0212 config at the size bench.py times (BASELINE.json configs[2]/[3]): 300^3-voxel main tensor
    (331x368x220, 1083 samples/ray), 160^3-voxel augmentation tensor, full 576x1024 frames."""
    configs, model_configs = H.load_configs(212, '00000')
    configs['model']['coarse_model']['num_voxels_initial'] = 300 ** 3
    configs['model']['coarse_model']['num_voxels_final'] = 300 ** 3
    configs['model']['augmentations'][0]['coarse_model']['num_voxels_initial'] = 160 ** 3
    configs['model']['augmentations'][0]['coarse_model']['num_voxels_final'] = 160 ** 3
    return configs, model_configs


def golden_tensorf_full_size():
    """The unmodified reference at the benchmarked size.  Per-ray maps, weights and the two masks are stored (bit-packed masks);
    the remaining per-sample tensors are re-derived by the oracle inside the GPU test (it is pinned bit-exact here)."""
    configs, model_configs = tensorf_full_size_configs()
    (OUT / 'tensorf_full_configs.json').write_text(json.dumps({'configs': configs, 'model_configs': model_configs}, indent=1))
    model = H.build_model(configs, model_configs)
    assert [int(v) for v in model.coarse_model.resolution] == [331, 368, 220] and int(model.coarse_model.num_samples) == 1083
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    sets = FX.tensorf_full_size_sets(configs, seed=31)
    load_tensorf_params(model, sets)
    keys = ('rgb', 'acc', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc', 'alpha', 'visibility', 'weights', 'raw_sigma', 'raw_rgb')
    for mode, R, seed in (('eval', 96, 7), ('train', 64, 8)):
        pixel_id = FX.random_pixels(R, nviews, h, w, seed)
        model.train(mode == 'train')
        torch.manual_seed(300 + seed)
        with torch.no_grad():
            ref = model({'pixel_id': pixel_id, 'num_frames': nviews, 'iter_num': 1, 'sub_batch_index': 1}, retraw=True)
        torch.manual_seed(300 + seed)
        with torch.no_grad():
            mine = P.tensorf_render_chunk(sets, configs, model_configs, pixel_id, training=(mode == 'train'))
        fixture = {'pixel_id': pixel_id, 'param_seed': 31, 'rng_seed': 300 + seed}
        prefixes = [''] + ([f"{a[0]}_" for a in sets['augmentations']] if mode == 'train' else [])
        for pre in prefixes:
            for k in keys:
                key = f'{pre}{k}_coarse'
                _check(f'tensorf_full/{mode}/{key}', ref[key], mine[key])
                if k in ('rgb', 'acc', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc', 'weights'):
                    fixture[key] = ref[key]
            for m in ('validity_mask', 'surface_mask'):
                fixture[f'{pre}{m}_coarse_bits'] = np.packbits(mine[f'{pre}{m}_coarse'].numpy().astype(np.uint8).reshape(-1))
        frac = mine['validity_mask_coarse'].float().mean().item(), mine['surface_mask_coarse'].float().mean().item()
        np.savez_compressed(OUT / f'tensorf_full_{mode}.npz', **_np(fixture))
        print(f'tensorf_full_{mode}: oracle == reference at 331x368x220 / 1083 samples (valid {frac[0]:.3f}, surface {frac[1]:.4f})')


def patch_loss_inputs(num_rays=6000, h=96, w=128, seed=0):
    """Seeded scene for the patch-reprojection losses: smooth random images, three LLFF-like cameras, rays through random
    pixels, two noisy depth candidates, and a half-random image-ray mask (shared with tests/)."""
    from simple_rf_b200 import synthetic
    mc = synthetic.scene_model_configs('llff', num_views=3)
    g = torch.Generator().manual_seed(seed)
    v = 3
    images = torch.rand(v, h, w, 3, generator=g)
    images = torch.nn.functional.avg_pool2d(images.permute(0, 3, 1, 2), 9, 1, 4).permute(0, 2, 3, 1).contiguous()
    poses = torch.tensor(np.array(mc['extrinsics']), dtype=torch.float32)
    k = torch.tensor(np.array(mc['intrinsics']), dtype=torch.float32)[0].clone()
    k[0, 2], k[1, 2], k[0, 0], k[1, 1] = w / 2, h / 2, 100.0, 100.0
    pid = torch.stack([torch.randint(0, v, (num_rays,), generator=g), torch.randint(0, w, (num_rays,), generator=g),
                       torch.randint(0, h, (num_rays,), generator=g)], 1)
    dirs = torch.stack([(pid[:, 1] - k[0, 2]) / k[0, 0], -(pid[:, 2] - k[1, 2]) / k[1, 1], -torch.ones(num_rays)], -1)
    c2w = poses[pid[:, 0]]
    rays_d = (c2w[:, :3, :3] @ dirs[..., None]).squeeze(-1)
    rays_o = c2w[:, :3, 3].contiguous()
    d1 = 2 + 6 * torch.rand(num_rays, generator=g)
    d2 = d1 * (1 + 0.2 * torch.randn(num_rays, generator=g))
    d2[::17] = -d2[::17]                                        # points behind the other camera
    d1[5::29] = 1e-3                                            # reprojections far outside the frame
    mask_nerf = torch.ones(num_rays, dtype=torch.bool)
    mask_nerf[num_rays // 2:] = torch.rand(num_rays - num_rays // 2, generator=g) < 0.5
    return dict(images=images, poses=poses, k=k, pixel_id=pid.int(), rays_o=rays_o, rays_d=rays_d, depth1=d1, depth2=d2,
                mask_nerf=mask_nerf, h=h, w=w)


def golden_patch_loss():
    """both reference loss classes' `compute_loss_nerf` on the seeded scene; the oracle must match
    loss, loss maps and gradients bit-exactly."""
    H.import_reference()
    from loss_functions.AugmentationsDepthLoss11 import AugmentationsDepthLoss
    from loss_functions.CoarseFineConsistencyLoss34 import CoarseFineConsistencyLoss
    a = patch_loss_inputs()
    fixture = {k_: v_ for k_, v_ in a.items()}
    n = a['depth1'].shape[0]
    intr = a['k'][None].expand(n, 3, 3)
    for tag, cls, rule in (('aug', AugmentationsDepthLoss, True), ('cf', CoarseFineConsistencyLoss, False)):
        d1 = a['depth1'].clone().requires_grad_()
        d2 = a['depth2'].clone().requires_grad_()
        obj = cls({'model': {'coarse_model': {}, 'fine_model': {}}, 'data_loader': {}}, {'patch_size': [5, 5], 'rmse_threshold': 0.1})
        loss, map1, map2 = obj.compute_loss_nerf(d1, d2, a['mask_nerf'], a['rays_o'], a['rays_d'], a['poses'], a['images'],
                                                 a['pixel_id'].long(), intr, (a['h'], a['w']))
        g1, g2 = torch.autograd.grad(loss, [d1, d2], allow_unused=True)
        g1 = torch.zeros(n) if g1 is None else g1
        g2 = torch.zeros(n) if g2 is None else g2
        m = a['mask_nerf']
        e1 = a['depth1'].clone().requires_grad_()
        e2 = a['depth2'].clone().requires_grad_()
        m1, m2, r1, r2 = OL.patch_reprojection_masks(a['rays_o'][m], a['rays_d'][m], e1[m], e2[m], a['pixel_id'][m], a['poses'], a['k'],
                                                     a['images'], (5, 5), 0.1, rule)
        oloss, omap1, omap2 = OL.masked_depth_loss(e1[m], e2[m], m1, m2)
        h1, h2 = torch.autograd.grad(oloss, [e1, e2], allow_unused=True)
        h1 = torch.zeros(n) if h1 is None else h1
        h2 = torch.zeros(n) if h2 is None else h2
        for name, ref, mine in (('loss', loss, oloss), ('map1', map1, omap1), ('map2', map2, omap2), ('g1', g1, h1), ('g2', g2, h2)):
            _check(f'patch_loss.{tag}.{name}', ref.detach(), mine.detach())
        fixture.update({f'{tag}_loss': loss, f'{tag}_map1': map1, f'{tag}_map2': map2, f'{tag}_g1': g1, f'{tag}_g2': g2,
                        f'{tag}_mask1': m1, f'{tag}_mask2': m2, f'{tag}_rmse1': r1, f'{tag}_rmse2': r2})
        print(f'patch_loss {tag}: oracle == reference (mask1 {m1.float().mean():.3f}, mask2 {m2.float().mean():.3f}, loss {loss.item():.5f})')
    np.savez_compressed(OUT / 'patch_loss.npz', **_np(fixture))


def golden_nerf_training_curve(iters=8, R=96):
    """Loss scalars per iteration (SURVEY.md §8c parity ledger): the UNMODIFIED reference model trained for a few iterations
    with the reference's optimiser (torch.optim.Adam over get_trainable_parameters, as OptimizerFactory02 builds it) on fixed
    random batches; the drop-in (forward + hand-written backward + fused Adam) must follow the same curve."""
    configs, model_configs = H.load_configs(1142, 'fern')
    model_configs = H.shrink(model_configs, 4)
    configs['model']['netchunk'] = 2048
    model = H.build_model(configs, model_configs)
    sets = FX.nerf_param_sets(configs, seed=11)
    load_nerf_params(model, sets)
    model.train()
    opt_cfg = next(c for c in configs['optimizers'] if c['name'] == 'optimizer_main')
    opt = torch.optim.Adam(model.get_trainable_parameters(opt_cfg), lr=opt_cfg['lr_initial'], betas=(opt_cfg['beta1'], opt_cfg['beta2']))
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    g = torch.Generator().manual_seed(77)
    pids, targets, losses = [], [], []
    torch.manual_seed(4242)
    rgb_keys = ('rgb_coarse', 'rgb_fine', 'points_augmentation_rgb_coarse', 'views_augmentation_rgb_coarse')
    for it in range(iters):
        if it % 4 == 0:                                # a new batch every 4 iterations: the loss falls within a batch
            pid = FX.random_pixels(R, nviews, h, w, 500 + it)
            target = torch.rand(R, 3, generator=g)
        opt.zero_grad(set_to_none=True)
        out = model({'pixel_id': pid, 'num_frames': nviews, 'iter_num': it, 'sub_batch_index': 0})
        loss = sum(((out[k] - target) ** 2).mean() for k in rgb_keys)
        loss = loss + 0.1 * (out['depth_coarse'] - out['points_augmentation_depth_coarse'].detach()).square().mean()
        loss.backward()
        opt.step()
        pids.append(pid); targets.append(target); losses.append(loss.detach())
        print(f'  reference training iteration {it}: loss {loss.item():.6f}')
    norms = torch.stack([p.detach().norm() for p in model.coarse_model.parameters()])
    fixture = {'pixel_id': torch.stack(pids), 'target': torch.stack(targets), 'loss': torch.stack(losses), 'param_seed': 11,
               'rng_seed': 4242, 'lr': opt_cfg['lr_initial'], 'beta1': opt_cfg['beta1'], 'beta2': opt_cfg['beta2'],
               'coarse_param_norms': norms}
    np.savez_compressed(OUT / 'nerf_train_curve.npz', **_np(fixture))
    print(f'nerf_train_curve: {iters} iterations of the reference model')


def tensorf_curve_configs():
    """Shipped TensoRF training config, shrunk for the CPU, with a resolution upsampling at iteration 4 and an alpha-mask
    rebuild (+ bounding-box shrink) at iteration 6 so that the curve crosses the in-forward model surgery and the
    optimiser re-grouping (src/models/SimpleTensoRF09.py:821-944)."""
    configs, model_configs = H.load_configs(212, '00000')
    model_configs = H.shrink(model_configs, 4)
    for cm, (v0, v1) in ((configs['model']['coarse_model'], (36, 52)), (configs['model']['augmentations'][0]['coarse_model'], (20, 28))):
        cm['num_voxels_initial'] = v0 ** 3
        cm['num_voxels_final'] = v1 ** 3
        cm['tensor_upsampling_iters'] = [4]
        cm['alpha_mask_update_iters'] = [6]
    return configs, model_configs


def golden_tensorf_training_curve(iters=8, R=96):
    """TensoRF counterpart of golden_nerf_training_curve: the unmodified reference model, its own parameter groups
    (lr_initial_tensor / lr_initial_network), torch.optim.Adam handed over through `model.optimizers` as Trainer10 does."""
    configs, model_configs = tensorf_curve_configs()
    (OUT / 'tensorf_curve_configs.json').write_text(json.dumps({'configs': configs, 'model_configs': model_configs}, indent=1))
    model = H.build_model(configs, model_configs)
    sets = FX.tensorf_sets(configs, seed=21, with_alpha=False)
    load_tensorf_params(model, sets)
    model.train()
    opt_cfg = next(c for c in configs['optimizers'] if c['name'] == 'optimizer_main')
    opt = torch.optim.Adam(model.get_trainable_parameters(opt_cfg), betas=(opt_cfg['beta1'], opt_cfg['beta2']))
    model.optimizers = {'optimizer_nerf': opt}
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    g = torch.Generator().manual_seed(78)
    pids, targets, losses, grids = [], [], [], []
    torch.manual_seed(4343)
    for it in range(iters):
        if it % 4 == 0:
            pid = FX.random_pixels(R, nviews, h, w, 600 + it)
            target = torch.rand(R, 3, generator=g)
        opt.zero_grad(set_to_none=True)
        out = model({'pixel_id': pid, 'num_frames': nviews, 'iter_num': it + 1, 'sub_batch_index': 0})
        loss = ((out['rgb_coarse'] - target) ** 2).mean() + ((out['points_augmentation_rgb_coarse'] - target) ** 2).mean()
        loss = loss + 0.1 * (out['depth_coarse'] - out['points_augmentation_depth_coarse'].detach()).square().mean()
        loss = loss + 1e-3 * out['weights_coarse'].square().sum(dim=1).mean()
        loss.backward()
        opt.step()
        pids.append(pid); targets.append(target); losses.append(loss.detach())
        grids.append(torch.as_tensor([int(v) for v in model.coarse_model.resolution]))
        print(f'  reference TensoRF iteration {it + 1}: loss {loss.item():.6f}, grid {grids[-1].tolist()}, '
              f'samples/ray {int(model.coarse_model.num_samples)}, alpha mask {model.coarse_model.alpha_mask is not None}')
    norms = torch.stack([p.detach().norm() for p in model.coarse_model.parameters()])
    fixture = {'pixel_id': torch.stack(pids), 'target': torch.stack(targets), 'loss': torch.stack(losses), 'param_seed': 21,
               'rng_seed': 4343, 'grid': torch.stack(grids), 'coarse_param_norms': norms,
               'bounding_box': model.coarse_model.bounding_box.detach().clone()}
    np.savez_compressed(OUT / 'tensorf_train_curve.npz', **_np(fixture))
    print(f'tensorf_train_curve: {iters} iterations of the reference model')


def golden_batch_assembly():
    """f2: the reference's own load_nerf_cached_batch / load_sparse_depth_cached_batch (unbound, on a stand-in object holding the
    cached tables) against oracle/batch.py, bit-exact; inputs + reference outputs -> tests/golden/batch_assembly.npz."""
    import types
    H.import_reference()
    from data_preprocessors.DataPreprocessor10 import DataPreprocessor
    from oracle import batch as OB
    t = OB.synthetic_tables()
    indices, m_nerf, m_sd = OB.synthetic_indices(t['pixel'].shape[0], 300, 212, seed=3)
    this = types.SimpleNamespace(device='cpu', preprocessed_data_dict={
        'frame_nums': np.arange(3), 'nerf_data': {'pixel_id': t['pixel'], 'target_rgb': t['rgb']},
        'sparse_depth_data': {'depths': t['depth'], 'reprojection_errors': t['error'], 'points_3d': t['points']}})
    idx = {'indices': indices, 'indices_mask_nerf': m_nerf, 'indices_mask_sparse_depth': m_sd}
    ref = DataPreprocessor.load_nerf_cached_batch(this, 7, idx)
    ref.update(DataPreprocessor.load_sparse_depth_cached_batch(this, idx, ref))
    mine = OB.assemble_batch(indices, m_nerf, m_sd, t['pixel'], t['rgb'], t['depth'], t['error'], t['points'])
    fixture = {'indices': indices, 'mask_nerf': m_nerf, 'mask_sd': m_sd}
    for k in ('pixel_id', 'target_rgb', 'sparse_depth_values', 'sparse_depth_errors', 'sparse_depth_points3d'):
        _check(f'batch/{k}', ref[k], mine[k])
        fixture[k] = ref[k]
    np.savez_compressed(OUT / 'batch_assembly.npz', **_np(fixture))
    print('batch_assembly: oracle == reference')


def surgery_configs(cp=False):
    """Shipped TensoRF config shrunk for the CPU: 48^3-voxel main tensor, one upsampling step to 64^3."""
    configs, model_configs = H.load_configs(212, '00000')
    model_configs = H.shrink(model_configs, 4)
    if cp:
        as_cp(configs)
    cm = configs['model']['coarse_model']
    cm['num_voxels_initial'], cm['num_voxels_final'] = 48 ** 3, 64 ** 3
    cm['tensor_upsampling_iters'] = [4]
    cm['alpha_mask_update_iters'] = [2, 6]
    configs['model']['augmentations'][0]['coarse_model']['num_voxels_initial'] = 20 ** 3
    return configs, model_configs


def _pack(volume):
    return np.packbits(volume.reshape(-1).numpy().astype(np.uint8))


def golden_surgery(cp=False):
    """f3: the reference's own update_alpha_mask / shrink_tensor / upsample_model_resolution on the seeded sparse tensor, in the
    order of its schedule (rebuild + shrink, upsample, rebuild against the previous mask at the old resolution); the oracle
    (oracle/surgery.py) must reproduce every volume, window, box and plane bit-exactly."""
    tag = 'tensorf_cp_surgery' if cp else 'tensorf_surgery'
    configs, model_configs = surgery_configs(cp)
    (OUT / f'{tag}_configs.json').write_text(json.dumps({'configs': configs, 'model_configs': model_configs}, indent=1))
    model = H.build_model(configs, model_configs)
    sets = FX.surgery_sets(configs, seed=41)
    load_tensorf_params(model, sets)
    t = model.coarse_model
    t.train()
    cfg = configs['model']['coarse_model']
    thr = cfg['alpha_mask_threshold']
    params = {k: v.clone() for k, v in sets['coarse_model']['params'].items()}
    geo = SG.tensor_geometry(sets['coarse_model']['resolution'], sets['coarse_model']['bbox'], cfg['num_voxels_per_sample'], cfg['num_samples_max'])
    fixture = {'param_seed': 41}
    with torch.no_grad():
        # 1. first rebuild (no previous mask) + shrink
        box_ref = t.update_alpha_mask(2)
        vol, box = SG.update_alpha_mask(params, geo, thr)
        _check('surgery/volume1', t.alpha_mask.alpha_volume[0, 0], vol)
        _check('surgery/box1', box_ref, box)
        fixture.update(volume1_bits=_pack(vol), volume1_shape=torch.tensor(vol.shape), box1=box_ref.clone(), occupied1=vol.mean())
        alpha_res = t.alpha_mask.resolution.clone()
        t.shrink_tensor(box_ref)
        t_l, b_r, box_s = SG.shrink_window(geo, box, alpha_res)
        params = SG.shrink_params(params, t_l, b_r)
        geo = SG.tensor_geometry(b_r - t_l, box_s, cfg['num_voxels_per_sample'], cfg['num_samples_max'])
        _check('surgery/shrink/resolution', t.resolution, geo['resolution'])
        _check('surgery/shrink/bbox', t.bounding_box, geo['bbox'])
        assert int(t.num_samples) == geo['num_samples']
        for k, v in dict(t.named_parameters()).items():
            if k.startswith(('matrices', 'vectors')):
                _check(f'surgery/shrink/{k}', v, params[k])
        fixture.update(window_lo=t_l, window_hi=b_r, shrink_resolution=t.resolution.clone(), shrink_bbox=t.bounding_box.clone(),
                       shrink_num_samples=int(t.num_samples))
        # 2. upsampling
        t.upsample_model_resolution(4)
        nv = SG.new_num_voxels(4, cfg['tensor_upsampling_iters'], cfg['num_voxels_initial'], cfg['num_voxels_final'])
        new_res = TF.vm_resolution(nv, geo['bbox'])
        params = SG.upsample_params(params, new_res)
        geo = SG.tensor_geometry(new_res, geo['bbox'], cfg['num_voxels_per_sample'], cfg['num_samples_max'])
        _check('surgery/upsample/resolution', t.resolution, geo['resolution'])
        for k, v in dict(t.named_parameters()).items():
            if k.startswith(('matrices', 'vectors')):
                _check(f'surgery/upsample/{k}', v, params[k])
        fixture.update(upsample_resolution=t.resolution.clone(), upsample_num_samples=int(t.num_samples),
                       upsampled_vectors_color_2=params['vectors_color.2'])
        if cp:
            fixture.update(upsampled_vectors_density_0=params['vectors_density.0'])
        else:
            fixture.update(upsampled_matrices_density_0=params['matrices_density.0'],
                           upsampled_matrices_color_1_sum=params['matrices_color.1'].double().sum())
        # 3. second rebuild: previous mask at the old resolution and the old (pre-shrink) box
        prev_vol, prev_box = t.alpha_mask.alpha_volume.clone(), t.alpha_mask.bounding_box.clone()
        box_ref2 = t.update_alpha_mask(6)
        vol2, box2 = SG.update_alpha_mask(params, geo, thr, prev_vol, prev_box)
        _check('surgery/volume2', t.alpha_mask.alpha_volume[0, 0], vol2)
        _check('surgery/box2', box_ref2, box2)
        fixture.update(volume2_bits=_pack(vol2), volume2_shape=torch.tensor(vol2.shape), box2=box_ref2.clone(), occupied2=vol2.mean())
    np.savez_compressed(OUT / f'{tag}.npz', **_np(fixture))
    print(f'{tag}: oracle == reference (grid {list(vol.shape)} {vol.mean():.3f} occupied -> window {t_l.tolist()}..{b_r.tolist()} '
          f'-> {geo["resolution"].tolist()}, second mask {vol2.mean():.3f} occupied)')


def _pose_probe(num_views, seed):
    """A small non-zero pose correction (axis-angle r, translation t) for every view."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(num_views, 3, generator=g) * 0.01, torch.randn(num_views, 3, generator=g) * 0.02


def _probe_loss(out, keys, seed):
    """A fixed random linear functional of the listed outputs (so that every one of them contributes to the pose gradient)."""
    g = torch.Generator().manual_seed(seed)
    return sum((out[k] * torch.randn(out[k].shape, generator=g)).sum() for k in keys)


def golden_learnable_cameras():
    """Learnable cameras (`learn_camera_rotation / learn_camera_translation`, SimpleNeRF17.py:817-842; no shipped run turns them on): the
    gradient of a fixed linear functional of the rendered maps w.r.t. the pose correction r, t, through the unmodified reference models in
    training mode (Simple-NeRF: NDC, coarse + fine + augmentations; Simple-TensoRF: NDC, main + augmentation tensor)."""
    from . import rays as RY
    fixtures = {}
    # ---- Simple-NeRF
    configs, model_configs = H.load_configs(1142, 'fern')
    model_configs = H.shrink(model_configs, 4)
    configs['model']['netchunk'] = 2048
    configs['model']['learn_camera_rotation'] = True
    configs['model']['learn_camera_translation'] = True
    model = H.build_model(configs, model_configs)
    sets = FX.nerf_param_sets(configs, seed=11)
    load_nerf_params(model, sets)
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    r0, t0 = _pose_probe(nviews, 31)
    model.extrinsics_learner.r.data.copy_(r0)
    model.extrinsics_learner.t.data.copy_(t0)
    pixel_id = FX.random_pixels(40, nviews, h, w, 9)
    keys = ('rgb_coarse', 'rgb_fine', 'depth_coarse', 'depth_fine', 'depth_ndc_fine', 'acc_fine', 'rays_o', 'rays_d_ndc', 'view_dirs')
    model.train(True)
    torch.manual_seed(909)
    ref = model({'pixel_id': pixel_id, 'num_frames': nviews, 'iter_num': 0, 'sub_batch_index': 0}, retraw=True)
    _probe_loss(ref, keys, 77).backward()
    g_r, g_t = model.extrinsics_learner.r.grad.clone(), model.extrinsics_learner.t.grad.clone()
    r1, t1 = r0.clone().requires_grad_(), t0.clone().requires_grad_()
    torch.manual_seed(909)
    mine = P.nerf_render_chunk(sets, configs, model_configs, pixel_id, training=True,
                               extrinsics=RY.pose_correction(torch.tensor(model_configs['extrinsics']), r1, t1))
    _probe_loss(mine, keys, 77).backward()
    _check('learnable_cameras/nerf/rgb_fine', ref['rgb_fine'].detach(), mine['rgb_fine'].detach(), exact=False, tol=2e-6)
    _check('learnable_cameras/nerf/r.grad', g_r, r1.grad, exact=False, tol=1e-4 * float(g_r.abs().max()))
    _check('learnable_cameras/nerf/t.grad', g_t, t1.grad, exact=False, tol=1e-4 * float(g_t.abs().max()))
    fixtures.update({'nerf_pixel_id': pixel_id, 'nerf_r': r0, 'nerf_t': t0, 'nerf_r_grad': g_r, 'nerf_t_grad': g_t,
                     'nerf_rgb_fine': ref['rgb_fine'].detach(), 'nerf_depth_fine': ref['depth_fine'].detach()})
    print(f'learnable cameras, Simple-NeRF: |r.grad| {float(g_r.norm()):.4f} |t.grad| {float(g_t.norm()):.4f}, oracle == reference')
    # ---- Simple-NeRF in world space (lindisp depths, white background): the rays enter through pts, view_dirs and |d| under delta
    configs, model_configs = nerf_variant_configs()
    configs['model']['learn_camera_rotation'] = True
    configs['model']['learn_camera_translation'] = True
    model = H.build_model(configs, model_configs)
    sets = FX.nerf_param_sets(configs, seed=13)
    load_nerf_params(model, sets)
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    r0, t0 = _pose_probe(nviews, 34)
    model.extrinsics_learner.r.data.copy_(r0)
    model.extrinsics_learner.t.data.copy_(t0)
    pixel_id = FX.random_pixels(40, nviews, h, w, 12)
    keys = ('rgb_coarse', 'rgb_fine', 'depth_coarse', 'depth_fine', 'depth_var_fine', 'acc_fine', 'rays_d', 'view_dirs')
    model.train(True)
    torch.manual_seed(912)
    ref = model({'pixel_id': pixel_id, 'num_frames': nviews, 'iter_num': 0, 'sub_batch_index': 0}, retraw=True)
    _probe_loss(ref, keys, 80).backward()
    g_r, g_t = model.extrinsics_learner.r.grad.clone(), model.extrinsics_learner.t.grad.clone()
    r1, t1 = r0.clone().requires_grad_(), t0.clone().requires_grad_()
    torch.manual_seed(912)
    mine = P.nerf_render_chunk(sets, configs, model_configs, pixel_id, training=True,
                               extrinsics=RY.pose_correction(torch.tensor(model_configs['extrinsics']), r1, t1))
    _probe_loss(mine, keys, 80).backward()
    _check('learnable_cameras/nerf_world/rgb_fine', ref['rgb_fine'].detach(), mine['rgb_fine'].detach(), exact=False, tol=2e-6)
    _check('learnable_cameras/nerf_world/r.grad', g_r, r1.grad, exact=False, tol=1e-4 * float(g_r.abs().max()))
    _check('learnable_cameras/nerf_world/t.grad', g_t, t1.grad, exact=False, tol=1e-4 * float(g_t.abs().max()))
    fixtures.update({'nerf_world_pixel_id': pixel_id, 'nerf_world_r': r0, 'nerf_world_t': t0, 'nerf_world_r_grad': g_r, 'nerf_world_t_grad': g_t,
                     'nerf_world_rgb_fine': ref['rgb_fine'].detach()})
    print(f'learnable cameras, Simple-NeRF world space: |r.grad| {float(g_r.norm()):.4f} |t.grad| {float(g_t.norm()):.4f}, oracle == reference')
    # ---- Simple-TensoRF (NDC)
    configs, model_configs = H.load_configs(212, '00000')
    model_configs = H.shrink(model_configs, 4)
    configs['model']['coarse_model']['num_voxels_initial'] = 40 ** 3
    configs['model']['augmentations'][0]['coarse_model']['num_voxels_initial'] = 20 ** 3
    configs['model']['learn_camera_rotation'] = True
    configs['model']['learn_camera_translation'] = True
    model = H.build_model(configs, model_configs)
    sets = FX.tensorf_sets(configs, seed=21, with_alpha=False)
    load_tensorf_params(model, sets)
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    r0, t0 = _pose_probe(nviews, 32)
    model.extrinsics_learner.r.data.copy_(r0)
    model.extrinsics_learner.t.data.copy_(t0)
    pixel_id = FX.random_pixels(32, nviews, h, w, 10)
    keys = ('rgb_coarse', 'depth_coarse', 'depth_ndc_coarse', 'acc_coarse', 'view_dirs', 'rays_d')
    model.train(True)
    torch.manual_seed(910)
    ref = model({'pixel_id': pixel_id, 'num_frames': nviews, 'iter_num': 1, 'sub_batch_index': 1}, retraw=True)
    _probe_loss(ref, keys, 78).backward()
    g_r, g_t = model.extrinsics_learner.r.grad.clone(), model.extrinsics_learner.t.grad.clone()
    r1, t1 = r0.clone().requires_grad_(), t0.clone().requires_grad_()
    torch.manual_seed(910)
    mine = P.tensorf_render_chunk(sets, configs, model_configs, pixel_id, training=True,
                                  extrinsics=RY.pose_correction(torch.tensor(model_configs['extrinsics']), r1, t1))
    _probe_loss(mine, keys, 78).backward()
    _check('learnable_cameras/tensorf/rgb', ref['rgb_coarse'].detach(), mine['rgb_coarse'].detach(), exact=False, tol=2e-6)
    _check('learnable_cameras/tensorf/r.grad', g_r, r1.grad, exact=False, tol=1e-4 * float(g_r.abs().max()))
    _check('learnable_cameras/tensorf/t.grad', g_t, t1.grad, exact=False, tol=1e-4 * float(g_t.abs().max()))
    fixtures.update({'tensorf_pixel_id': pixel_id, 'tensorf_r': r0, 'tensorf_t': t0, 'tensorf_r_grad': g_r, 'tensorf_t_grad': g_t,
                     'tensorf_rgb': ref['rgb_coarse'].detach(), 'tensorf_depth': ref['depth_coarse'].detach()})
    print(f'learnable cameras, Simple-TensoRF: |r.grad| {float(g_r.norm()):.4f} |t.grad| {float(g_t.norm()):.4f}, oracle == reference')
    # ---- Simple-TensoRF in world space: the box-march depths start at the ray's entry into the box, so they move with the pose (:388-400)
    configs, model_configs = tensorf_world_configs()
    configs['model']['learn_camera_rotation'] = True
    configs['model']['learn_camera_translation'] = True
    model = H.build_model(configs, model_configs)
    sets = FX.tensorf_sets(configs, seed=23, with_alpha=False)
    load_tensorf_params(model, sets)
    h, w = model_configs['resolution']
    nviews = len(model_configs['intrinsics'])
    r0, t0 = _pose_probe(nviews, 33)
    model.extrinsics_learner.r.data.copy_(r0)
    model.extrinsics_learner.t.data.copy_(t0)
    pixel_id = FX.random_pixels(32, nviews, h, w, 11)
    keys = ('rgb_coarse', 'depth_coarse', 'depth_var_coarse', 'acc_coarse', 'view_dirs', 'rays_d', 'z_vals_coarse')
    model.train(True)
    torch.manual_seed(911)
    ref = model({'pixel_id': pixel_id, 'num_frames': nviews, 'iter_num': 1, 'sub_batch_index': 1}, retraw=True)
    _probe_loss(ref, keys, 79).backward()
    g_r, g_t = model.extrinsics_learner.r.grad.clone(), model.extrinsics_learner.t.grad.clone()
    r1, t1 = r0.clone().requires_grad_(), t0.clone().requires_grad_()
    torch.manual_seed(911)
    mine = P.tensorf_render_chunk(sets, configs, model_configs, pixel_id, training=True,
                                  extrinsics=RY.pose_correction(torch.tensor(model_configs['extrinsics']), r1, t1))
    _probe_loss(mine, keys, 79).backward()
    _check('learnable_cameras/tensorf_world/rgb', ref['rgb_coarse'].detach(), mine['rgb_coarse'].detach(), exact=False, tol=2e-6)
    _check('learnable_cameras/tensorf_world/r.grad', g_r, r1.grad, exact=False, tol=1e-4 * float(g_r.abs().max()))
    _check('learnable_cameras/tensorf_world/t.grad', g_t, t1.grad, exact=False, tol=1e-4 * float(g_t.abs().max()))
    fixtures.update({'tensorf_world_pixel_id': pixel_id, 'tensorf_world_r': r0, 'tensorf_world_t': t0, 'tensorf_world_r_grad': g_r,
                     'tensorf_world_t_grad': g_t, 'tensorf_world_rgb': ref['rgb_coarse'].detach(),
                     'tensorf_world_depth': ref['depth_coarse'].detach()})
    print(f'learnable cameras, Simple-TensoRF world space: |r.grad| {float(g_r.norm()):.4f} |t.grad| {float(g_t.norm()):.4f}, oracle == reference')
    np.savez_compressed(OUT / 'learnable_cameras.npz', **_np(fixtures))


def main():
    if not H.available():
        sys.exit('reference checkout not available: goldens can only be regenerated in the build container')
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    if 'learnable_cameras' in sys.argv[1:]:
        golden_learnable_cameras()
        return
    if 'nerf_variants' in sys.argv[1:]:
        golden_nerf_variants()
        return
    if 'tensorf_variants' in sys.argv[1:]:
        golden_tensorf_variants()
        return
    if 'cp' in sys.argv[1:]:                     # only the CANDECOMP/PARAFAC fixtures
        golden_tensorf_cp()
        golden_surgery(cp=True)
        return
    configs, model_configs = golden_nerf()
    golden_sample_pdf()
    golden_composite(copy.deepcopy(configs), model_configs)
    golden_tensorf()
    if 'full' in sys.argv[1:] or not (OUT / 'tensorf_full_eval.npz').exists():
        golden_tensorf_full_size()
    golden_patch_loss()
    golden_nerf_training_curve()
    golden_tensorf_training_curve()
    golden_batch_assembly()
    golden_surgery()
    golden_tensorf_world()
    golden_tensorf_cp()
    golden_surgery(cp=True)
    golden_nerf_variants()
    golden_tensorf_variants()
    golden_learnable_cameras()


if __name__ == '__main__':
    main()
