"""Oracle: Simple-TensoRF grid surgery (test infrastructure; "next" row f3 of SURVEY.md §8f).

CPU restatement, in functional torch ops on parameter dicts, of the occasional model modifications of
src/models/SimpleTensoRF09.py: alpha-mask rebuild (:849-876 with compute_alpha :878-897), bounding-box shrink (:899-914,
:1299-1320), resolution schedule (:837-847) and plane / line upsampling (:1284-1297).  Pinned bit-exact against the
unmodified reference by oracle/generate_golden.py::golden_surgery (tests/golden/tensorf_surgery.npz).
"""
import numpy
import torch
import torch.nn.functional as F

from . import tensorf as TF


def tensor_geometry(resolution, bbox, voxels_per_sample=0.5, num_samples_max=1e6):
    """SimpleTensoRF09.py:652-665 (update_tensor_params)."""
    bbox = bbox.float()
    size = bbox[1] - bbox[0]
    resolution = resolution.long()
    voxel_length = size / (resolution - 1)
    return {'bbox': bbox, 'size': size, 'resolution': resolution, 'voxel_length': voxel_length,
            'step_size': torch.mean(voxel_length) * voxels_per_sample,
            'num_samples': TF.vm_num_samples(resolution, voxels_per_sample, num_samples_max)}


def dense_grid(bbox, resolution):
    """SimpleTensoRF09.py:850-855: world coordinates of the grid points, [X,Y,Z,3]."""
    gs = [int(v) for v in resolution]
    samples = torch.stack(torch.meshgrid(torch.linspace(0, 1, gs[0]), torch.linspace(0, 1, gs[1]), torch.linspace(0, 1, gs[2]),
                                         indexing='ij'), -1)
    return bbox[0] * (1 - samples) + bbox[1] * samples


def compute_alpha(params, bbox, xyz, length, prev_volume=None, prev_bbox=None, density_predictor='ReLU', density_offset=-10.0):
    """SimpleTensoRF09.py:878-897 at world points xyz [N,3]."""
    if prev_volume is not None:
        live = TF.sample_alpha(prev_volume, prev_bbox, xyz) > 0
    else:
        live = torch.ones_like(xyz[:, 0], dtype=bool)
    sigma = torch.zeros(xyz.shape[:-1])
    if live.any():
        pn = TF.normalize(xyz, bbox)
        sigma = TF.density(params, pn, live, density_predictor, density_offset)
    return 1 - torch.exp(-sigma * length).view(xyz.shape[:-1])


def update_alpha_mask(params, geometry, threshold, prev_volume=None, prev_bbox=None, density_predictor='ReLU', density_offset=-10.0):
    """SimpleTensoRF09.py:849-876 -> (alpha volume float {0,1} [Z,Y,X], new bounding box [2,3])."""
    gs = tuple(int(v) for v in geometry['resolution'])
    dense_xyz = dense_grid(geometry['bbox'], gs)
    alpha = compute_alpha(params, geometry['bbox'], dense_xyz.view(-1, 3), geometry['step_size'], prev_volume, prev_bbox,
                          density_predictor, density_offset).view(gs)
    dense_xyz = dense_xyz.transpose(0, 2).contiguous()
    alpha = alpha.clamp(0, 1).transpose(0, 2).contiguous()[None, None]
    alpha = F.max_pool3d(alpha, kernel_size=3, padding=1, stride=1).view(gs[::-1])
    alpha = (alpha >= threshold).float()
    valid = dense_xyz[alpha > 0.5]
    return alpha, torch.stack((valid.amin(0), valid.amax(0)))


def shrink_window(geometry, new_bbox, alpha_resolution):
    """SimpleTensoRF09.py:899-914 -> (t_l, b_r voxel windows [3] long, bounding box after the correction of :905-910)."""
    lo, hi = new_bbox
    bbox, res, voxel = geometry['bbox'], geometry['resolution'], geometry['voxel_length']
    t_l, b_r = (lo - bbox[0]) / voxel, (hi - bbox[0]) / voxel
    t_l, b_r = torch.round(torch.round(t_l)).long(), torch.round(b_r).long() + 1
    b_r = torch.stack([b_r, res]).amin(0)
    if not torch.equal(torch.as_tensor(alpha_resolution).long(), res):
        t_l_r, b_r_r = t_l / (res - 1), (b_r - 1) / (res - 1)
        box = torch.zeros_like(new_bbox)
        box[0] = (1 - t_l_r) * bbox[0] + t_l_r * bbox[1]
        box[1] = (1 - b_r_r) * bbox[0] + b_r_r * bbox[1]
        new_bbox = box
    return t_l, b_r, new_bbox


def shrink_params(params, t_l, b_r):
    """SimpleTensoRF09.py:1299-1320 (VM) / :1113-1124 (CP): window slices of every plane / line."""
    out = dict(params)
    for kind in ('density', 'color'):
        for i in range(3):
            v = TF.VECTOR_AXES[i]
            a0, a1 = TF.MATRIX_AXES[i]
            out[f'vectors_{kind}.{i}'] = params[f'vectors_{kind}.{i}'][..., t_l[v]:b_r[v], :]
            if f'matrices_{kind}.{i}' in params:
                out[f'matrices_{kind}.{i}'] = params[f'matrices_{kind}.{i}'][..., t_l[a1]:b_r[a1], t_l[a0]:b_r[a0]]
    return out


def new_num_voxels(iter_num, upsampling_iters, num_voxels_initial, num_voxels_final):
    """SimpleTensoRF09.py:837-847: log-linear voxel-count ladder."""
    if iter_num not in upsampling_iters:
        raise RuntimeError('get_new_num_voxels() called at invalid iteration number')
    k = upsampling_iters.index(iter_num) + 1
    lo, hi = numpy.log(num_voxels_initial), numpy.log(num_voxels_final)
    return int(numpy.round(numpy.exp(lo + (hi - lo) * k / len(upsampling_iters))))


def upsample_params(params, new_resolution):
    """SimpleTensoRF09.py:1284-1297 (VM) / :1101-1111 (CP): bilinear, align_corners=True."""
    out = dict(params)
    res = [int(v) for v in new_resolution]
    for kind in ('density', 'color'):
        for i in range(3):
            a0, a1 = TF.MATRIX_AXES[i]
            v = TF.VECTOR_AXES[i]
            if f'matrices_{kind}.{i}' in params:
                out[f'matrices_{kind}.{i}'] = F.interpolate(params[f'matrices_{kind}.{i}'], size=(res[a1], res[a0]), mode='bilinear', align_corners=True)
            out[f'vectors_{kind}.{i}'] = F.interpolate(params[f'vectors_{kind}.{i}'], size=(res[v], 1), mode='bilinear', align_corners=True)
    return out


def reconfigure_optimizer(optimizer, fresh_groups, log=None):
    """SimpleTensoRF09.py:916-944 on a torch optimiser: per fresh group, walk the existing groups (deleting from the list
    being enumerated, as the reference does), drop `num_params` state entries at the running ordinal position, then append the
    fresh groups.  `log` collects the 'unable to delete' events the reference prints."""
    for fresh in fresh_groups:
        position = 0
        for i, existing in enumerate(optimizer.param_groups):
            n = len(existing['params'])
            if existing['name'] == fresh['name']:
                del optimizer.param_groups[i]
                for _ in range(n):
                    keys = list(optimizer.state.keys())
                    if position < len(keys):
                        del optimizer.state[keys[position]]
                    elif log is not None:
                        log.append((fresh['name'], position, len(keys)))
            else:
                position += n
    for fresh in fresh_groups:
        optimizer.add_param_group(fresh)


def pack_volume(volume):
    return torch.from_numpy(numpy.packbits(volume.reshape(-1).numpy().astype(numpy.uint8)))


def replay_golden_schedule(configs):
    """The golden's sequence: rebuild + crop at 2, resample at 4, rebuild against the old mask at 6."""
    cfg = configs['model']['coarse_model']
    from . import fixtures as FX
    t = FX.surgery_sets(configs, seed=41)['coarse_model']
    params = {k: v.clone() for k, v in t['params'].items()}
    geo = tensor_geometry(t['resolution'], t['bbox'], cfg['num_voxels_per_sample'], cfg['num_samples_max'])
    out = {'geo0': geo, 'params0': params}
    vol1, box1 = update_alpha_mask(params, geo, cfg['alpha_mask_threshold'])
    lo, hi, box_s = shrink_window(geo, box1, torch.tensor([vol1.shape[2], vol1.shape[1], vol1.shape[0]]))
    params = shrink_params(params, lo, hi)
    geo = tensor_geometry(hi - lo, box_s, cfg['num_voxels_per_sample'], cfg['num_samples_max'])
    out.update(vol1=vol1, box1=box1, lo=lo, hi=hi, geo1=geo, params1=params, prev_box=out['geo0']['bbox'])
    nv = new_num_voxels(4, cfg['tensor_upsampling_iters'], cfg['num_voxels_initial'], cfg['num_voxels_final'])
    res = TF.vm_resolution(nv, geo['bbox'])
    params = upsample_params(params, res)
    geo = tensor_geometry(res, geo['bbox'], cfg['num_voxels_per_sample'], cfg['num_samples_max'])
    out.update(geo2=geo, params2=params)
    vol2, box2 = update_alpha_mask(params, geo, cfg['alpha_mask_threshold'], vol1[None, None], out['geo0']['bbox'])
    out.update(vol2=vol2, box2=box2)
    return out
