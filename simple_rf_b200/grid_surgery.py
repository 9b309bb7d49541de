"""Grid surgery of a Simple-TensoRF VM tensor ("next" row f3) on the kernels of csrc/tensorf_surgery.cu.

What the reference does inside `LowRankTensor.run_model_modifications` (src/models/SimpleTensoRF09.py:821-830) on five of its
25 000 iterations, organised here as a precomputed plan per tensor:

    plan[iter] = [rebuild occupancy (+ crop to the occupied box on the first rebuild)] + [resample to the next voxel count]

* occupancy rebuild (:849-876, compute_alpha :878-897): `srf_alpha_grid_occupancy` evaluates the density of every voxel
  straight into one bit (no fp32 [Z,Y,X] volume, no per-point arrays), `srf_alpha_grid_dilate` does the 3x3x3 max-pool as a
  bit dilation and returns the new bool volume plus the occupied set's per-axis projections; the new bounding box is read off
  the same per-axis coordinate arrays the kernel evaluated (amin / amax of :872-873).
* crop (:899-914, :1299-1320) and resample (:1284-1297): `srf_resample_plane` writes contiguous new parameters (the reference
  keeps strided views of the old storage after the crop; the values are identical).
* optimiser re-grouping (:916-944): same observable result as the reference, including its ordinal-position state deletion
  and the re-added groups starting from the INITIAL learning rates.
"""
import ctypes
import warnings

import numpy
import torch

from . import _lib as L
from . import tensorf_ops as T


# ---------------------------------------------------------------------------------------------------- plan
def voxel_ladder(tensor_configs):
    """iteration -> voxel count after that iteration's upsampling (:837-847): log-linear between initial and final."""
    iters = list(tensor_configs['tensor_upsampling_iters'])
    lo, hi = numpy.log(tensor_configs['num_voxels_initial']), numpy.log(tensor_configs['num_voxels_final'])
    return {it: int(numpy.round(numpy.exp(lo + (hi - lo) * (k + 1) / len(iters)))) for k, it in enumerate(iters)}


def build_plan(tensor_configs):
    """iteration -> tuple of steps, in the reference's order (:822-829)."""
    ladder = voxel_ladder(tensor_configs)
    rebuilds = list(tensor_configs['alpha_mask_update_iters'])
    plan = {}
    for it in sorted(set(rebuilds) | set(ladder)):
        steps = []
        if it in rebuilds:
            steps.append(('occupancy', it == rebuilds[0]))          # crop only with the first rebuild (:824)
        if it in ladder:
            steps.append(('resample', ladder[it]))
        plan[it] = tuple(steps)
    return plan


# ---------------------------------------------------------------------------------------------------- occupancy
def axis_coordinates(bounding_box, resolution):
    """World coordinates of the grid planes along each axis: the 1-D factors of the dense grid of :850-855 (linspace on the
    host, blend on the device, un-fused — the same bits as the reference's [X,Y,Z,3] tensor)."""
    dev = bounding_box.device
    out = []
    for a in range(3):
        s = torch.linspace(0, 1, int(resolution[a])).to(dev)
        out.append((bounding_box[0, a] * (1 - s) + bounding_box[1, a] * s).contiguous())
    return out


@torch.no_grad()
def rebuild_occupancy(planes, lines, geometry, *, step_size, threshold, softplus, density_offset, previous=None):
    """-> (bool volume [Z,Y,X], new bounding box [2,3]).  planes / lines: the density parameters ([1,C,H,W] / [1,C,L,1]; planes None:
    a CP tensor);
    geometry: dict(box [2,3] tensor, box_min, box_size, res) of the tensor; previous: AlphaGridMask.packed() or None."""
    box = geometry['box']
    dev = box.device
    res = [int(v) for v in geometry['res']]
    if planes is None:                       # CANDECOMP/PARAFAC tensor: the three lines only (:1043-1062)
        cl_planes, cl_lines = None, T.lines_channels_last(list(lines))
    else:
        cl_planes, cl_lines = T.to_channels_last(list(planes), list(lines))
    coords = axis_coordinates(box, res)
    c_res = T._i3(res)
    words = L.load().srf_alpha_grid_words(c_res)
    raw = torch.empty((words,), dtype=torch.int32, device=dev)
    channels = (ctypes.c_int * 3)(*[l.shape[-1] for l in cl_lines])
    if previous is None:
        p_bits = p_res = p_min = p_size = None
    else:
        p_bits, p_res, p_min, p_size = L.ptr(previous['bits']), T._i3(previous['res']), T._f3(previous['box_min']), T._f3(previous['box_size'])
    n_vox = res[0] * res[1] * res[2]
    L.call('srf_alpha_grid_occupancy', T._ptrs(cl_planes) if cl_planes is not None else None, T._ptrs(cl_lines), channels, c_res, T._f3(geometry['box_min']),
           T._f3(geometry['box_size']), L.ptr(coords[0]), L.ptr(coords[1]), L.ptr(coords[2]), p_bits, p_res, p_min, p_size,
           int(bool(softplus)), float(density_offset), float(step_size), float(threshold), L.ptr(raw), L.stream_handle(),
           work=float(n_vox) * (576.0 if cl_planes is not None else 4.0 * 6 * cl_lines[0].shape[-1]))     # requested texel bytes per voxel
    volume = torch.empty((res[2], res[1], res[0]), dtype=torch.uint8, device=dev)
    pitch = (res[0] + 31) // 32
    projection = torch.zeros((pitch + res[1] + res[2],), dtype=torch.int32, device=dev)
    L.call('srf_alpha_grid_dilate', L.ptr(raw), c_res, L.ptr(volume), L.ptr(projection), L.stream_handle(),
           work=float(n_vox) * (1.0 + 9.0 / 8.0))
    proj = projection.cpu().numpy()                                   # the one synchronisation of a rebuild
    x_words = proj[:pitch].astype(numpy.uint32)
    x_flags = ((x_words[:, None] >> numpy.arange(32, dtype=numpy.uint32)[None]) & 1).reshape(-1)[:res[0]].astype(bool)
    flags = [x_flags, proj[pitch:pitch + res[1]] != 0, proj[pitch + res[1]:] != 0]
    if not all(f.any() for f in flags):
        raise RuntimeError('alpha-mask rebuild: no voxel reaches alpha_mask_threshold (the reference fails at amin of an empty tensor, SimpleTensoRF09.py:872)')
    corners = [[], []]
    for a in range(3):
        occupied = coords[a][torch.from_numpy(flags[a]).to(dev)]
        corners[0].append(occupied.amin())
        corners[1].append(occupied.amax())
    return volume.view(torch.bool), torch.stack([torch.stack(corners[0]), torch.stack(corners[1])])


# ---------------------------------------------------------------------------------------------------- crop / resample
@torch.no_grad()
def resample(param, out_hw, window=None):
    """New contiguous [1,C,oh,ow] tensor: bilinear (align_corners=True) resampling of `window` = (y0, x0, h, w) of
    param [1,C,H,W] (default: all of it); an output extent equal to the window is a copy."""
    L.require_cuda(param)
    src = L.f32c(param.detach())
    _, C, H, W = src.shape
    y0, x0, h, w = window if window is not None else (0, 0, H, W)
    oh, ow = int(out_hw[0]), int(out_hw[1])
    dst = torch.empty((1, C, oh, ow), dtype=torch.float32, device=src.device)
    L.call('srf_resample_plane', L.ptr(src), C, H, W, int(y0), int(x0), int(h), int(w), L.ptr(dst), oh, ow, L.stream_handle(),
           work=float(dst.numel()) * 4 + float(C * h * w) * 4)
    return dst


def crop_window(bounding_box, voxel_length, resolution, new_box, alpha_resolution):
    """Voxel window [lo, hi) per axis and the bounding box that goes with it (:899-914).  Host arithmetic on fp32 torch scalars
    with the reference's operations (the double rounding of `lo` included), so the window and box bits are the same."""
    box, voxel, res = bounding_box.cpu(), voxel_length.cpu(), resolution.cpu()
    new_box = new_box.cpu()
    lo = torch.round(torch.round((new_box[0] - box[0]) / voxel)).long()
    hi = torch.minimum(torch.round((new_box[1] - box[0]) / voxel).long() + 1, res)
    if not torch.equal(alpha_resolution.cpu(), res):
        # the mask was built on another grid: snap the box to this grid's voxel centres (:905-910)
        f_lo, f_hi = lo / (res - 1), (hi - 1) / (res - 1)
        new_box = torch.stack([(1 - f_lo) * box[0] + f_lo * box[1], (1 - f_hi) * box[0] + f_hi * box[1]])
    return lo, hi, new_box


def map_vm_parameters(matrices, vectors, fn_plane, fn_line):
    """Apply fn_plane(param, axis0, axis1) / fn_line(param, axis) to the three plane / line parameters -> ParameterLists."""
    mats = [torch.nn.Parameter(fn_plane(matrices[i], *T.MATRIX_AXES[i])) for i in range(3)]
    vecs = [torch.nn.Parameter(fn_line(vectors[i], T.VECTOR_AXES[i])) for i in range(3)]
    return torch.nn.ParameterList(mats), torch.nn.ParameterList(vecs)


def crop_vm(matrices, vectors, lo, hi):
    lo, hi = [int(v) for v in lo], [int(v) for v in hi]
    return map_vm_parameters(
        matrices, vectors,
        lambda m, a0, a1: resample(m, (hi[a1] - lo[a1], hi[a0] - lo[a0]), (lo[a1], lo[a0], hi[a1] - lo[a1], hi[a0] - lo[a0])),
        lambda v, a: resample(v, (hi[a] - lo[a], 1), (lo[a], 0, hi[a] - lo[a], 1)))


def resample_vm(matrices, vectors, new_res):
    res = [int(v) for v in new_res]
    return map_vm_parameters(matrices, vectors, lambda m, a0, a1: resample(m, (res[a1], res[a0])), lambda v, a: resample(v, (res[a], 1)))


def crop_lines(vectors, lo, hi):
    """CP tensor (:1113-1124): the window of each line."""
    lo, hi = [int(v) for v in lo], [int(v) for v in hi]
    return torch.nn.ParameterList([torch.nn.Parameter(resample(vectors[i], (hi[a] - lo[a], 1), (lo[a], 0, hi[a] - lo[a], 1)))
                                   for i, a in enumerate(T.VECTOR_AXES)])


def resample_lines(vectors, new_res):
    """CP tensor (:1101-1111)."""
    res = [int(v) for v in new_res]
    return torch.nn.ParameterList([torch.nn.Parameter(resample(vectors[i], (res[a], 1))) for i, a in enumerate(T.VECTOR_AXES)])


# ---------------------------------------------------------------------------------------------------- optimiser
def regroup_optimizer(optimizer, fresh_groups):
    """Swap this tensor's parameter groups for `fresh_groups` (same names).  Observable behaviour of :916-944: the old group is
    removed, as many optimiser-state entries as it had parameters are dropped BY ORDINAL POSITION (the position of the group's
    first parameter among all parameters of the groups in front of it — not by parameter identity, so entries of other
    parameters go when some parameter never received a state), and the fresh groups are appended with the learning rates of the
    optimiser configuration (not the decayed ones)."""
    for fresh in fresh_groups:
        groups = optimizer.param_groups
        at = next((i for i, g in enumerate(groups) if g.get('name') == fresh['name']), None)
        if at is None:
            continue
        first = sum(len(g['params']) for g in groups[:at])
        count = len(groups[at]['params'])
        del groups[at]
        doomed = list(optimizer.state.keys())[first:first + count]
        for key in doomed:
            del optimizer.state[key]
        if len(doomed) < count:
            warnings.warn(f"{fresh['name']}: only {len(doomed)} of {count} optimiser-state entries existed at position {first}")
    for fresh in fresh_groups:
        optimizer.add_param_group(fresh)
