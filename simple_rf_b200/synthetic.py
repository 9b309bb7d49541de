"""Synthetic LLFF- / RealEstate-shaped scenes and config dicts for benchmarks and smoke tests.

Shapes follow the reference's shipped runs (runs/training/train1142/Configs.json for Simple-NeRF,
runs/training/train0212/Configs.json for Simple-TensoRF): LLFF frames are 756x1008 with
focal 815.13, RealEstate frames 576x1024 with focal 493.91; cameras are forward-facing with small
baselines, NDC sampling, near plane 1.  No dataset or checkpoint is needed: throughput of the render
path does not depend on the weights, which are default-initialised.
"""
import copy
import math

import numpy as np

_NERF_MLP = {
    'points_net_depth': 8, 'views_net_depth': 1, 'points_net_width': 256, 'views_net_width': 128,
    'points_positional_encoding_degree': 10, 'views_positional_encoding_degree': 4,
    'use_view_dirs': True, 'view_dependent_rgb': True, 'predict_visibility': False,
}


def nerf_configs(model_name='SimpleNeRF91', augmentations=True, rng_mode='reference'):
    coarse = dict(_NERF_MLP, num_samples=64)
    fine = dict(_NERF_MLP, num_samples=128)
    model = {
        'name': model_name, 'coarse_model': coarse, 'fine_model': fine,
        'learn_camera_focal_length': False, 'learn_camera_rotation': False, 'learn_camera_translation': False,
        'chunk': 4096, 'lindisp': False, 'netchunk': 16384, 'perturb': True, 'raw_noise_std': 1.0,
        'white_bkgd': False, 'rng_mode': rng_mode,
    }
    if augmentations:
        pa = dict(_NERF_MLP, points_sigma_positional_encoding_degree=3)
        va = dict(_NERF_MLP, use_view_dirs=False, view_dependent_rgb=False)
        del va['views_positional_encoding_degree']
        va['views_positional_encoding_degree'] = 4
        model['augmentations'] = [{'name': 'points_augmentation', 'coarse_model': pa},
                                  {'name': 'views_augmentation', 'coarse_model': va}]
    return {
        'database': 'NeRF_LLFF', 'data_loader': {'ndc': True, 'num_rays': 2048, 'sparse_depth': {'num_rays': 2048}},
        'model': model, 'sub_batch_size': 4096, 'seed': 0, 'device': [0],
        'optimizers': [{'name': 'optimizer_main', 'beta1': 0.9, 'beta2': 0.999, 'lr_initial': 5e-4, 'lr_decay': 250}],
    }


_VM_TENSOR = {
    'decomposition_type': 'VectorMatrix', 'num_samples_max': 1e6, 'num_components_density': [16, 4, 4],
    'num_components_color': [48, 12, 12], 'bounding_box': [[-1.5, -1.67, -1.0], [1.5, 1.67, 1.0]],
    'num_voxels_initial': 128 ** 3 + 4, 'num_voxels_final': 300 ** 3, 'tensor_upsampling_iters': [2000, 3000, 4000, 5500],
    'num_voxels_per_sample': 0.5, 'alpha_mask_update_iters': [2500], 'alpha_mask_threshold': 1e-4,
    'ray_marching_weight_threshold': 1e-4, 'use_view_dirs': True, 'view_dependent_color': True,
    'views_positional_encoding_degree': 0, 'features_positional_encoding_degree': 0, 'features_dimension_color': 27,
    'density_offset': -10, 'distance_scale': 25, 'density_predictor': 'ReLU', 'color_predictor': 'MLP_Features',
    'num_units_color_predictor': 128, 'predict_visibility': False,
}


def tensorf_configs(model_name='SimpleTensoRF91', num_voxels=None, augmentations=True, rng_mode='reference'):
    main = copy.deepcopy(_VM_TENSOR)
    if num_voxels is not None:
        main['num_voxels_initial'] = num_voxels
    model = {
        'name': model_name, 'coarse_model': main,
        'learn_camera_focal_length': False, 'learn_camera_rotation': False, 'learn_camera_translation': False,
        'chunk': 4096, 'lindisp': False, 'perturb': True, 'white_bkgd': False, 'rng_mode': rng_mode,
    }
    if augmentations:
        aug = copy.deepcopy(_VM_TENSOR)
        aug.update(num_components_density=[4, 4, 4], bounding_box=[[-1.5, -1.67, -0.5], [1.5, 1.67, 1.0]],
                   num_voxels_initial=64 ** 3, num_voxels_final=160 ** 3)
        model['augmentations'] = [{'name': 'points_augmentation', 'coarse_model': aug}]
    return {
        'database': 'RealEstate10K', 'data_loader': {'ndc': True, 'num_rays': 2048, 'sparse_depth': {'num_rays': 2048}},
        'model': model, 'sub_batch_size': 4096, 'seed': 0, 'device': [0],
        'optimizers': [{'name': 'optimizer_main', 'beta1': 0.9, 'beta2': 0.99, 'lr_initial_tensor': 0.02,
                        'lr_initial_network': 0.001}],
    }


def _pose(tx, ty, tz, yaw):
    c, s = math.cos(yaw), math.sin(yaw)
    m = np.eye(4, dtype=np.float64)
    m[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    m[:3, 3] = [tx, ty, tz]
    return m


def scene_model_configs(kind='llff', num_views=3):
    """`model_configs` dict with the keys of DataPreprocessor10.get_model_configs (:70-88)."""
    if kind == 'llff':
        h, w, f = 756, 1008, 815.1316
        bounds = (1.0, 8.0)
    else:
        h, w, f = 576, 1024, 493.9102
        bounds = (1.0, 100.0)
    k = [[f, 0.0, w / 2.0], [0.0, f, h / 2.0], [0.0, 0.0, 1.0]]
    poses = [_pose(0.25 * (i - (num_views - 1) / 2.0), 0.03 * ((i % 2) * 2 - 1), 0.0, 0.04 * (i - (num_views - 1) / 2.0))
             for i in range(num_views)]
    return {
        'resolution': [h, w], 'bounds': list(bounds), 'translation_scale': 1.0,
        'train_frame_nums': list(range(num_views)),
        'intrinsics': [copy.deepcopy(k) for _ in range(num_views)],
        'extrinsics': [p.tolist() for p in poses],
        'near': 1.0, 'far': float(bounds[1]), 'near_ndc': 0.0, 'far_ndc': 1.0,
        'bounding_box': [[-1.5, -1.67, -1.0], [1.5, 1.67, 1.0]],
    }


def trajectory_pose(model_configs, t):
    """A test pose on a small spiral around the training cameras, t in [0,1)."""
    a = 2 * math.pi * t
    return _pose(0.3 * math.cos(a), 0.05 * math.sin(a), 0.05 * math.sin(2 * a), 0.03 * math.cos(a))


def frame_pixel_ids(height, width, view=0):
    """All (view, x, y) triples of one frame in the row-major order of create_test_data
    (DataPreprocessor10.py:736-743)."""
    ys, xs = np.meshgrid(np.arange(height, dtype=np.int32), np.arange(width, dtype=np.int32), indexing='ij')
    return np.stack([np.full(xs.size, view, dtype=np.int32), xs.reshape(-1), ys.reshape(-1)], axis=1)


def blocky_alpha_volume(size, block, p_block, p_speckle, generator):
    """[Z,Y,X] float {0,1} occupancy volume from platform-independent ops only (mt19937 `rand`, comparisons, exact
    nearest-neighbour repetition): blocks of `block`^3 voxels occupied with probability p_block + isolated voxels with
    probability p_speckle — contiguous occupied regions (what a trained scene gives) with a ragged surface."""
    import torch
    n = -(-size // block)
    coarse = torch.rand(n, n, n, generator=generator) < p_block
    vol = coarse.repeat_interleave(block, 0).repeat_interleave(block, 1).repeat_interleave(block, 2)[:size, :size, :size]
    vol = vol | (torch.rand(size, size, size, generator=generator) < p_speckle)
    return vol.float()
