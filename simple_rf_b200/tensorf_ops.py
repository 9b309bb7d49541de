"""Tensor-level wrappers + autograd glue for the Simple-TensoRF kernels (csrc/tensorf.cu).

Reference: models/SimpleTensoRF09.py:701-761 (LowRankTensor.forward), :1214-1272 (VM density / colour),
:1342-1349 (AlphaGridMask).  No CPU / eager fallback.
"""
import ctypes

import torch

from . import _lib as L

MATRIX_AXES = [[0, 1], [0, 2], [1, 2]]      # SimpleTensoRF09.py:1131
VECTOR_AXES = [2, 1, 0]                     # SimpleTensoRF09.py:1132


def _f3(t):
    v = [float(x) for x in torch.as_tensor(t).reshape(-1).tolist()]
    return (ctypes.c_float * len(v))(*v)


def _i3(t):
    v = [int(x) for x in torch.as_tensor(t).reshape(-1).tolist()]
    return (ctypes.c_int * len(v))(*v)


def _ptrs(tensors):
    return (ctypes.c_void_p * len(tensors))(*[L.ptr(t) for t in tensors])


def pack_alpha_bits(alpha_volume):
    """{0,1} volume [..., Z, Y, X] (AlphaGridMask.alpha_volume: bool / uint8 in the drop-in, fp32 in the reference's memory
    form) -> uint32 words, 1 bit per voxel, x fastest."""
    L.require_cuda(alpha_volume)
    n = alpha_volume.numel()
    bits = torch.empty(((n + 31) // 32,), dtype=torch.int32, device=alpha_volume.device)
    if alpha_volume.dtype in (torch.bool, torch.uint8):
        vol = alpha_volume.contiguous().view(torch.uint8).reshape(-1)
        L.call('srf_pack_alpha_bits_u8', L.ptr(vol), n, L.ptr(bits), L.stream_handle())
    else:
        vol = L.f32c(alpha_volume).reshape(-1)
        L.call('srf_pack_alpha_bits', L.ptr(vol), n, L.ptr(bits), L.stream_handle())
    return bits


def corner_or_alpha_bits(bits, res):
    """Derived cache of the fused march: per cell of the alpha grid the OR of its 8 corner bits, (X+1)(Y+1)(Z+1) bits."""
    c_res = _i3(res)
    out = torch.empty((L.load().srf_alpha_corner_or_words(c_res),), dtype=torch.int32, device=bits.device)
    L.call('srf_alpha_corner_or_bits', L.ptr(bits), c_res, L.ptr(out), L.stream_handle())
    return out


class Compacted:
    """A compacted sample list: int32 indices (first `count` valid) with the count kept on the device."""
    __slots__ = ('mask', 'idx', 'count', 'total')

    def __init__(self, mask, idx, count, total):
        self.mask, self.idx, self.count, self.total = mask, idx, count, total


def _compact(mask_u8, counts, total):
    dev = mask_u8.device
    offsets = torch.empty_like(counts)
    idx = torch.empty((max(total, 1),), dtype=torch.int32, device=dev)
    count = torch.empty((1,), dtype=torch.int32, device=dev)
    L.call('srf_compact', L.ptr(mask_u8), total, L.ptr(counts), L.ptr(offsets), L.ptr(idx), L.ptr(count), L.stream_handle())
    return idx, count


def validity_compact(rays_o, rays_d, z, bbox, alpha=None):
    """bbox test AND alphaMask test on pts = o + d z, then stable compaction (SimpleTensoRF09.py:705-710, :1221).
    alpha: None or dict(bits=int32 words, res=(X,Y,Z), box_min=[3], box_size=[3])."""
    L.require_cuda(rays_o, rays_d, z)
    rays_o, rays_d, z = L.f32c(rays_o), L.f32c(rays_d), L.f32c(z)
    R, S = z.shape
    total = R * S
    mask = torch.empty((R, S), dtype=torch.uint8, device=z.device)
    nb = L.load().srf_compaction_blocks(total)
    counts = torch.empty((max(nb, 1),), dtype=torch.int32, device=z.device)
    if alpha is None:
        a_bits = a_res = a_min = a_size = None
    else:
        a_bits, a_res, a_min, a_size = L.ptr(alpha['bits']), _i3(alpha['res']), _f3(alpha['box_min']), _f3(alpha['box_size'])
    L.call('srf_tensorf_mask', L.ptr(rays_o), L.ptr(rays_d), L.ptr(z), R, S, _f3(bbox), a_bits, a_res, a_min, a_size,
           L.ptr(mask), L.ptr(counts), L.stream_handle(), work=float(total) * 5)      # 4 B depth in + 1 B mask out per sample
    idx, count = _compact(mask, counts, total)
    return Compacted(mask.view(torch.bool), idx, count, total)


def threshold_compact(values, threshold):
    """mask = values > threshold, then stable compaction (SimpleTensoRF09.py:726, :1248)."""
    L.require_cuda(values)
    values = L.f32c(values.detach())
    total = values.numel()
    mask = torch.empty(values.shape, dtype=torch.uint8, device=values.device)
    nb = L.load().srf_compaction_blocks(total)
    counts = torch.empty((max(nb, 1),), dtype=torch.int32, device=values.device)
    L.call('srf_threshold_mask', L.ptr(values), float(threshold), total, L.ptr(mask), L.ptr(counts), L.stream_handle())
    idx, count = _compact(mask, counts, total)
    return Compacted(mask.view(torch.bool), idx, count, total)


_CL_CACHE = {}


def to_channels_last(planes, lines):
    """[1,C,H,W] -> [H,W,C] and [1,C,L,1] -> [L,C] contiguous copies (the derived caches the kernels read).  Rebuilt only
    when a parameter's storage or version changes, so the chunks of one frame (and all frames of a test run) share them."""
    key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in (*planes, *lines))
    slot = (len(planes), planes[0].device, planes[0].shape[1], lines[0].shape[1] if lines else 0, id(planes[0]))
    hit = _CL_CACHE.get(slot)
    if hit is not None and hit[0] == key:
        return hit[1]
    val = ([p.detach()[0].permute(1, 2, 0).contiguous() for p in planes],
           [l.detach()[0, :, :, 0].permute(1, 0).contiguous() for l in lines])
    if len(_CL_CACHE) > 64:
        _CL_CACHE.clear()
    _CL_CACHE[slot] = (key, val)
    return val


class VmGeometry:
    """Everything the gather kernels need besides the parameter tables."""

    def __init__(self, rays_o, rays_d, z, box_min, box_size, resolution):
        """z: [R,S] per-ray depths, or a [S] ladder shared by all rays (test time)."""
        self.rays_o, self.rays_d, self.z = L.f32c(rays_o), L.f32c(rays_d), L.f32c(z)
        self.z_is_ladder = z.dim() == 1
        self.S = z.shape[-1]
        self.box_min, self.box_size = _f3(box_min), _f3(box_size)
        self.res = _i3(resolution)

    def args(self, comp):
        return (L.ptr(self.rays_o), L.ptr(self.rays_d), L.ptr(self.z), self.S, L.ptr(comp.idx), L.ptr(comp.count), comp.total,
                self.box_min, self.box_size)


class _VmDensity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, geom, comp, softplus, offset, n_planes, *params):
        planes, lines = params[:n_planes], params[n_planes:]
        planes_cl, lines_cl = to_channels_last(planes, lines)
        chans = _i3([p.shape[1] for p in planes])
        R = geom.z.shape[0]
        sigma = torch.zeros((R, geom.S, 1), dtype=torch.float32, device=geom.z.device)
        feat = torch.empty((max(comp.total, 1),), dtype=torch.float32, device=geom.z.device)
        L.call('srf_vm_density_fwd', *geom.args(comp), _ptrs(planes_cl), _ptrs(lines_cl), chans, geom.res, int(softplus),
               float(offset), L.ptr(sigma), L.ptr(feat), L.stream_handle(),
               work=(comp.count, 4.0 * 6 * sum(p.shape[1] for p in planes)))        # requested texel bytes: 4 plane + 2 line texels x C floats
        ctx.geom, ctx.comp, ctx.cfg = geom, comp, (int(softplus), float(offset), n_planes)
        ctx.tables = (planes_cl, lines_cl, chans)
        ctx.save_for_backward(feat)
        return sigma

    @staticmethod
    def backward(ctx, g_sigma):
        (feat,) = ctx.saved_tensors
        geom, comp = ctx.geom, ctx.comp
        softplus, offset, n_planes = ctx.cfg
        planes_cl, lines_cl, chans = ctx.tables
        gp = [torch.zeros_like(p) for p in planes_cl]
        gl = [torch.zeros_like(l) for l in lines_cl]
        g = L.f32c(g_sigma)                      # bound to a local: a converted copy must outlive the launch call
        L.call('srf_vm_density_bwd', *geom.args(comp), _ptrs(planes_cl), _ptrs(lines_cl), chans, geom.res, softplus, offset,
               L.ptr(g), L.ptr(feat), _ptrs(gp), _ptrs(gl), L.stream_handle(),
               work=(comp.count, 2 * 4.0 * 6 * sum(p.shape[2] for p in planes_cl)))   # texel reads + the same again as atomic adds
        grads = [g.permute(2, 0, 1)[None].contiguous() for g in gp] + [g.permute(1, 0)[None, :, :, None].contiguous() for g in gl]
        return (None, None, None, None, None, *grads)


def vm_density(geom, comp, planes, lines, softplus=False, offset=0.0):
    """sigma [R,S,1] (zeros where the mask is false), differentiable w.r.t. planes / lines ([1,C,H,W] / [1,C,L,1])."""
    return _VmDensity.apply(geom, comp, softplus, offset, len(planes), *planes, *lines)


def color_row_pitch(num_products):
    """bf16 elements per colour row: products + 3 view directions, rounded up to one 16-wide MMA K step."""
    return -(-(num_products + 3) // 16) * 16


def vm_color_rows(geom, comp, view_dirs, planes, lines, max_rows=None):
    """rows [total, pitch] bf16 = [plane x line products (sum C) | view_dirs (3) | 0], first comp.count rows valid; also
    returns the channels-last tables for vm_color_rows_backward.  Not an autograd node by itself: the colour branch
    (gather -> MLP) is one autograd.Function in models/SimpleTensoRF91.py so the bf16 rows never carry a gradient."""
    planes_cl, lines_cl = to_channels_last(planes, lines)
    chans = _i3([p.shape[1] for p in planes])
    pitch = color_row_pitch(sum(p.shape[1] for p in planes))
    rows = torch.empty((max(comp.total if max_rows is None else max_rows, 1), pitch), dtype=torch.bfloat16, device=geom.z.device)
    vd = L.f32c(view_dirs)
    L.call('srf_vm_color_features_fwd', *geom.args(comp), _ptrs(planes_cl), _ptrs(lines_cl), chans, geom.res,
           L.ptr(vd), L.ptr(rows), pitch, int(geom.z_is_ladder), L.stream_handle(),
           work=(comp.count, 4.0 * 6 * sum(p.shape[1] for p in planes)))
    return rows, (planes_cl, lines_cl, chans)


def vm_color_rows_backward(geom, comp, tables, g_rows):
    """g_rows [>= count, >= sum C] fp32 -> gradients of planes ([1,C,H,W]) and lines ([1,C,L,1])."""
    planes_cl, lines_cl, chans = tables
    gp = [torch.zeros_like(p) for p in planes_cl]
    gl = [torch.zeros_like(l) for l in lines_cl]
    g = L.f32c(g_rows)
    L.call('srf_vm_color_features_bwd', *geom.args(comp), _ptrs(planes_cl), _ptrs(lines_cl), chans, geom.res,
           L.ptr(g), g.shape[1], _ptrs(gp), _ptrs(gl), L.stream_handle(),
           work=(comp.count, 2 * 4.0 * 6 * sum(p.shape[2] for p in planes_cl)))
    return ([x.permute(2, 0, 1)[None].contiguous() for x in gp], [x.permute(1, 0)[None, :, :, None].contiguous() for x in gl])


# ---------------------------------------------------------------------------------------------------- CANDECOMP/PARAFAC tensor
def lines_channels_last(lines):
    """[1,C,L,1] -> [L,C] derived caches of a CP tensor's three lines (to_channels_last without planes)."""
    key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in lines)
    slot = ('cp', lines[0].device, lines[0].shape[1], id(lines[0]))
    hit = _CL_CACHE.get(slot)
    if hit is not None and hit[0] == key:
        return hit[1]
    val = [l.detach()[0, :, :, 0].permute(1, 0).contiguous() for l in lines]
    if len(_CL_CACHE) > 64:
        _CL_CACHE.clear()
    _CL_CACHE[slot] = (key, val)
    return val


class _CpDensity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, geom, comp, softplus, offset, *lines):
        lines_cl = lines_channels_last(lines)
        C = lines[0].shape[1]
        R = geom.z.shape[0]
        sigma = torch.zeros((R, geom.S, 1), dtype=torch.float32, device=geom.z.device)
        feat = torch.empty((max(comp.total, 1),), dtype=torch.float32, device=geom.z.device)
        L.call('srf_cp_density_fwd', *geom.args(comp), _ptrs(lines_cl), C, geom.res, int(softplus), float(offset), L.ptr(sigma),
               L.ptr(feat), L.stream_handle(), work=(comp.count, 4.0 * 6 * C))           # requested bytes: 6 line texels x C floats
        ctx.geom, ctx.comp, ctx.cfg, ctx.tables = geom, comp, (int(softplus), float(offset), C), lines_cl
        ctx.save_for_backward(feat)
        return sigma

    @staticmethod
    def backward(ctx, g_sigma):
        (feat,) = ctx.saved_tensors
        geom, comp, lines_cl = ctx.geom, ctx.comp, ctx.tables
        softplus, offset, C = ctx.cfg
        gl = [torch.zeros_like(l) for l in lines_cl]
        g = L.f32c(g_sigma)
        L.call('srf_cp_density_bwd', *geom.args(comp), _ptrs(lines_cl), C, geom.res, softplus, offset, L.ptr(g), L.ptr(feat), _ptrs(gl),
               L.stream_handle(), work=(comp.count, 2 * 4.0 * 6 * C))
        return (None, None, None, None, *[x.permute(1, 0)[None, :, :, None].contiguous() for x in gl])


def cp_density(geom, comp, lines, softplus=False, offset=0.0):
    """sigma [R,S,1] (zeros where the mask is false) of a CP tensor, differentiable w.r.t. the three lines [1,C,L,1]."""
    return _CpDensity.apply(geom, comp, softplus, offset, *lines)


def cp_color_rows(geom, comp, view_dirs, lines, max_rows=None):
    """rows [total, pitch] bf16 = [line products (C) | view_dirs (3) | 0] of a CP tensor (the A operand of the colour MLP) + the
    channels-last tables for cp_color_rows_backward."""
    lines_cl = lines_channels_last(lines)
    C = lines[0].shape[1]
    pitch = color_row_pitch(C)
    rows = torch.empty((max(comp.total if max_rows is None else max_rows, 1), pitch), dtype=torch.bfloat16, device=geom.z.device)
    vd = L.f32c(view_dirs)
    L.call('srf_cp_color_features_fwd', *geom.args(comp), _ptrs(lines_cl), C, geom.res, L.ptr(vd), L.ptr(rows), pitch, L.stream_handle(),
           work=(comp.count, 4.0 * 6 * C))
    return rows, (lines_cl, C)


def cp_color_rows_backward(geom, comp, tables, g_rows):
    """g_rows [>= count, >= C] fp32 -> gradients of the three lines ([1,C,L,1])."""
    lines_cl, C = tables
    gl = [torch.zeros_like(l) for l in lines_cl]
    g = L.f32c(g_rows)
    L.call('srf_cp_color_features_bwd', *geom.args(comp), _ptrs(lines_cl), C, geom.res, L.ptr(g), g.shape[1], _ptrs(gl), L.stream_handle(),
           work=(comp.count, 2 * 4.0 * 6 * C))
    return [x.permute(1, 0)[None, :, :, None].contiguous() for x in gl]


class Marched:
    """Result of the fused test-time march: per-ray maps + the flat surface list (indices in (ray, sample) order, their
    weights, per-ray offsets / counts; the total stays on the device)."""
    __slots__ = ('maps', 'surface', 'weights', 'ray_offset', 'ray_count')


def march(rays_o_ndc, rays_d_ndc, rays_o, rays_d, ladder, bbox, box_size, alpha, planes, lines, resolution, *, softplus, offset,
          distance_scale, threshold, use_corner_or=True):
    """Fused test-time pass of one VM tensor (csrc/tensorf_march.cu): box + alphaMask test, density, transmittance, per-ray maps
    and the surface list, with no [R,S] intermediate.  ladder [S]: the sample depths shared by all rays."""
    L.require_cuda(rays_o_ndc, rays_d_ndc, rays_o, rays_d, ladder)
    rays_o_ndc, rays_d_ndc, rays_o, rays_d, ladder = (L.f32c(t) for t in (rays_o_ndc, rays_d_ndc, rays_o, rays_d, ladder))
    R, S = rays_o_ndc.shape[0], ladder.shape[0]
    dev = ladder.device
    planes_cl, lines_cl = to_channels_last(planes, lines)
    chans = _i3([p.shape[1] for p in planes])
    maps = {k: torch.empty((R,), dtype=torch.float32, device=dev) for k in ('acc', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc')}
    ray_count = torch.empty((R,), dtype=torch.int32, device=dev)
    entry_sample = torch.empty((R, S), dtype=torch.int32, device=dev)
    entry_weight = torch.empty((R, S), dtype=torch.float32, device=dev)
    if alpha is None:
        a_bits = a_coarse = a_res = a_min = a_size = None
    else:
        a_bits, a_res, a_min, a_size = L.ptr(alpha['bits']), _i3(alpha['res']), _f3(alpha['box_min']), _f3(alpha['box_size'])
        a_coarse = None
        if use_corner_or:                        # cached next to the bits it was derived from
            if 'corner_or' not in alpha:
                alpha['corner_or'] = corner_or_alpha_bits(alpha['bits'], alpha['res'])
            a_coarse = L.ptr(alpha['corner_or'])
    L.call('srf_tensorf_march', L.ptr(rays_o_ndc), L.ptr(rays_d_ndc), L.ptr(rays_o), L.ptr(rays_d), L.ptr(ladder), R, S, _f3(bbox),
           _f3(box_size), a_bits, a_coarse, a_res, a_min, a_size, _ptrs(planes_cl), _ptrs(lines_cl), chans, _i3(resolution), int(softplus),
           float(offset), float(distance_scale), float(threshold), *[L.ptr(maps[k]) for k in ('acc', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc')],
           L.ptr(ray_count), L.ptr(entry_sample), L.ptr(entry_weight), L.stream_handle(), work=float(R) * S * 4)
    nb = L.load().srf_tensorf_march_blocks(R)
    scratch = torch.empty((R + nb,), dtype=torch.int32, device=dev)
    ray_offset = torch.empty((R,), dtype=torch.int32, device=dev)
    idx = torch.empty((max(R * S, 1),), dtype=torch.int32, device=dev)
    weights = torch.empty((max(R * S, 1),), dtype=torch.float32, device=dev)
    count = torch.empty((1,), dtype=torch.int32, device=dev)
    L.call('srf_tensorf_march_compact', L.ptr(ray_count), R, S, L.ptr(entry_sample), L.ptr(entry_weight), L.ptr(scratch), L.ptr(ray_offset),
           L.ptr(idx), L.ptr(weights), L.ptr(count), L.stream_handle())
    m = Marched()
    m.maps, m.surface, m.weights, m.ray_offset, m.ray_count = maps, Compacted(None, idx, count, R * S), weights, ray_offset, ray_count
    return m


def ray_accumulate(rgb_rows, marched, white_bkgd):
    """rgb_map [R,3] = sum of weight * colour over every ray's surface samples (+ 1 - acc on a white background)."""
    R = marched.ray_count.shape[0]
    rgb_map = torch.empty((R, 3), dtype=torch.float32, device=rgb_rows.device)
    rgb_rows = L.f32c(rgb_rows)
    L.call('srf_ray_accumulate', L.ptr(rgb_rows), L.ptr(marched.weights), L.ptr(marched.ray_offset), L.ptr(marched.ray_count),
           L.ptr(marched.maps['acc']), R, int(bool(white_bkgd)), L.ptr(rgb_map), L.stream_handle())
    return rgb_map


def scatter_rows(comp, src, width, total):
    dst = torch.zeros((total, width), dtype=torch.float32, device=src.device)
    src = L.f32c(src)
    L.call('srf_scatter_rows', L.ptr(comp.idx), L.ptr(comp.count), comp.total, L.ptr(src), width, L.ptr(dst), L.stream_handle())
    return dst


def gather_rows(comp, src, width, nrows=None):
    """nrows: rows of the result when the caller knows count <= nrows (default: the worst case comp.total)."""
    dst = torch.zeros((max(comp.total if nrows is None else nrows, 1), width), dtype=torch.float32, device=src.device)
    src = L.f32c(src)
    L.call('srf_gather_rows', L.ptr(comp.idx), L.ptr(comp.count), comp.total, L.ptr(src), width, L.ptr(dst), L.stream_handle())
    return dst


class _ScatterRows(torch.autograd.Function):
    """dense[idx[j]] = rows[j] for j < count (the reference's rgb[mask] = surface_rgb, SimpleTensoRF09.py:1271)."""

    @staticmethod
    def forward(ctx, comp, rows, total):
        ctx.comp, ctx.width, ctx.nrows = comp, rows.shape[1], rows.shape[0]
        return scatter_rows(comp, rows, rows.shape[1], total)

    @staticmethod
    def backward(ctx, g):
        return None, gather_rows(ctx.comp, g, ctx.width, ctx.nrows)[:ctx.nrows], None
