"""Seeded synthetic parameters / pixels shared by oracle/generate_golden.py and tests/ (test infrastructure).

Model parameters are never stored in fixtures: both sides regenerate them from a seeded
torch.Generator (CPU mt19937 streams are platform independent)."""
import torch

from . import nerf_mlp as M
from . import tensorf as TF


def nerf_param_sets(configs, seed):
    g = torch.Generator().manual_seed(seed)
    mc = configs['model']
    sets = {'coarse_model': M.init_mlp_params(mc['coarse_model'], g),
            'fine_model': M.init_mlp_params(mc['fine_model'], g), 'augmentations': []}
    for aug in mc.get('augmentations', []):
        sets['augmentations'].append((aug['name'], aug['coarse_model'], M.init_mlp_params(aug['coarse_model'], g)))
    # a sigma bias keeps the random-init field from being empty (weights/depths exercise the scan)
    for p in [sets['coarse_model'], sets['fine_model']] + [a[2] for a in sets['augmentations']]:
        p['pts_output_linear.bias'][0] += 2.0
    return sets


def random_pixels(n, num_views, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randint(0, num_views, (n,), generator=g),
                        torch.randint(0, w, (n,), generator=g),
                        torch.randint(0, h, (n,), generator=g)], 1).int()


def tensorf_sets(configs, seed, with_alpha):
    g = torch.Generator().manual_seed(seed)
    mc = configs['model']

    def one(cfg):
        bbox = torch.tensor(cfg['bounding_box'])
        res = TF.vm_resolution(cfg['num_voxels_initial'], bbox)
        cp = cfg['decomposition_type'] == 'CandecompParafac'
        init = TF.init_cp_params if cp else TF.init_vm_params
        t = {'params': init(res, cfg['num_components_density'], cfg['num_components_color'], generator=g, use_views=cfg['use_view_dirs']),
             'bbox': bbox, 'resolution': res, 'num_samples': TF.vm_num_samples(res, cfg['num_voxels_per_sample'], cfg['num_samples_max'])}
        # random-init planes give sigma ~ 0; scale density up so weights cross the 1e-4 surface threshold
        # (CP: the feature is a sum of products of THREE 0.1 randn factors)
        for i in range(3):
            if cp:
                t['params'][f'vectors_density.{i}'] *= 4.5
                t['params'][f'vectors_color.{i}'] *= 2.2          # products of the size the VM fixture has (0.1 x 0.1)
            else:
                t['params'][f'matrices_density.{i}'] *= 6.0
        if with_alpha:
            X, Y, Z = [int(r) for r in res]
            vol = (torch.rand(Z, Y, X, generator=g) < 0.35).float()
            t['alpha_volume'] = vol.view(1, 1, Z, Y, X)
            t['alpha_bbox'] = bbox.clone()
        return t
    sets = {'coarse_model': one(mc['coarse_model']), 'augmentations': []}
    for aug in mc.get('augmentations', []):
        sets['augmentations'].append((aug['name'], aug['coarse_model'], one(aug['coarse_model'])))
    return sets




from simple_rf_b200.synthetic import blocky_alpha_volume  # noqa: E402  (shared with bench.py: same mask in tests and bench)


def tensorf_full_size_sets(configs, seed, alpha_size=190):
    """BASELINE.json configs[2]/[3] size: the tensors of `configs` at their configured `num_voxels_initial` (300^3 ->
    331x368x220 voxels, 1083 samples/ray for the main tensor) with a 190^3 alpha mask on the main tensor — the shapes bench.py
    times.  Same recipe as tensorf_sets(): seeded 0.1 randn planes, density planes x6."""
    g = torch.Generator().manual_seed(seed)
    mc = configs['model']

    def one(cfg, with_alpha):
        bbox = torch.tensor(cfg['bounding_box'])
        res = TF.vm_resolution(cfg['num_voxels_initial'], bbox)
        t = {'params': TF.init_vm_params(res, cfg['num_components_density'], cfg['num_components_color'], generator=g),
             'bbox': bbox, 'resolution': res, 'num_samples': TF.vm_num_samples(res, cfg['num_voxels_per_sample'], cfg['num_samples_max'])}
        for i in range(3):
            t['params'][f'matrices_density.{i}'] *= 6.0
        if with_alpha:
            vol = blocky_alpha_volume(alpha_size, 10, 0.10, 0.004, g)
            t['alpha_volume'] = vol.view(1, 1, alpha_size, alpha_size, alpha_size)
            t['alpha_bbox'] = bbox.clone()
        return t
    sets = {'coarse_model': one(mc['coarse_model'], True), 'augmentations': []}
    for aug in mc.get('augmentations', []):
        sets['augmentations'].append((aug['name'], aug['coarse_model'], one(aug['coarse_model'], False)))
    return sets


def sparsify_density(t, floor=0.6):
    """Random-init planes put sigma above the alpha-mask threshold almost everywhere (after the 3^3 pooling: everywhere).  A
    constant negative density component (channel 0 of plane 0 = -floor, of line 0 = 1) leaves a few per cent of the voxels
    occupied, so the rebuilt mask, its dilation and the shrunk bounding box are non-trivial.  (CP: component 0 of the three
    lines = -floor, 1, 1.)"""
    if TF.is_cp(t['params']):
        t['params']['vectors_density.0'][:, 0] = -floor
        t['params']['vectors_density.1'][:, 0] = 1.0
        t['params']['vectors_density.2'][:, 0] = 1.0
        return t
    t['params']['matrices_density.0'][:, 0] = -floor
    t['params']['vectors_density.0'][:, 0] = 1.0
    return t


def carve_empty_border(t, margins=((0.15, 0.1), (0.1, 0.2), (0.2, 0.12))):
    """Push the density far below zero near the faces of the box (per axis: fraction of the extent at the low / high side),
    so the occupied voxels do not touch the box and shrink_tensor has something to cut.  (CP: component 1 + a carries the ramp
    of axis a in its line, -50 and 1 in the other two lines.)"""
    res = [int(r) for r in t['resolution']]
    for a, (lo, hi) in enumerate(margins):
        n = res[a]
        ramp = torch.zeros(n)
        ramp[: int(lo * n)] = 1.0
        ramp[n - int(hi * n):] = 1.0
        i = TF.VECTOR_AXES.index(a)                      # the line running along axis a
        if TF.is_cp(t['params']):
            j, k = [q for q in range(3) if q != i]
            t['params'][f'vectors_density.{i}'][0, 1 + a, :, 0] = ramp
            t['params'][f'vectors_density.{j}'][0, 1 + a, :, 0] = -50.0
            t['params'][f'vectors_density.{k}'][0, 1 + a, :, 0] = 1.0
            continue
        t['params'][f'vectors_density.{i}'][0, 1, :, 0] = ramp
        t['params'][f'matrices_density.{i}'][0, 1] = -50.0
    return t


def surgery_sets(configs, seed=41):
    """Tensors for the grid-surgery fixtures (alpha-mask rebuild, shrink, upsampling): tensorf_sets() + sparse density."""
    sets = tensorf_sets(configs, seed, with_alpha=False)
    carve_empty_border(sparsify_density(sets['coarse_model']))
    return sets
