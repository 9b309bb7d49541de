"""Tensor-level wrappers over the C ABI (device pointers in, tensors out) + autograd glue.

Every function launches hand-written sm_100a kernels on the current CUDA stream; none has a
CPU / eager fallback.
"""
import torch

from . import _lib as L


def _empty(shape, like, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device=like.device)


# ------------------------------------------------------------------------------------------------
# rays (reference: utils/CommonUtils04.py:73-149)
# ------------------------------------------------------------------------------------------------
def camera_tables(intrinsics, c2w, device):
    """Per-view tables the raygen kernel reads: K^-1 (torch.linalg.inv on the F host matrices, as the
    reference inverts per-ray copies of them, CommonUtils04.py:87), c2w, (fx, fy)."""
    k = torch.as_tensor(intrinsics, dtype=torch.float32).detach().cpu()
    e = torch.as_tensor(c2w, dtype=torch.float32).detach().cpu()
    k_inv = torch.linalg.inv(k).reshape(-1, 9)
    focal = torch.stack([k[:, 0, 0], k[:, 1, 1]], 1)
    return (k_inv.contiguous().to(device), e.reshape(-1, 16).contiguous().to(device), focal.contiguous().to(device))


def raygen(pixel_id, tables, height, width, near, *, half_pixel, flip_x, ndc, viewdirs_from_ndc):
    L.require_cuda(pixel_id)
    k_inv, c2w, focal = tables
    pid = pixel_id.to(torch.int32).contiguous()
    R = pid.shape[0]
    outs = [_empty((R, 3), pid) for _ in range(5)]
    near = float(near)
    near_f = torch.tensor(near, dtype=torch.float32).item()
    two_near_f = torch.tensor(2.0 * near, dtype=torch.float32).item()
    L.call('srf_raygen', L.ptr(pid), R, L.ptr(k_inv), L.ptr(c2w), L.ptr(focal), k_inv.shape[0], int(height), int(width),
           near_f, two_near_f, int(half_pixel), int(flip_x), int(ndc), int(viewdirs_from_ndc),
           *[L.ptr(o) for o in outs], L.stream_handle())
    rays_o, rays_d, o_ndc, d_ndc, vd = outs
    if not ndc:
        o_ndc = d_ndc = None
    return rays_o, rays_d, o_ndc, d_ndc, vd


# ------------------------------------------------------------------------------------------------
# sampling (reference: models/SimpleNeRF17.py:330-417)
# ------------------------------------------------------------------------------------------------
def stratified_z(ladder, num_rays, jitter=None, philox_seed=None):
    """ladder [S] device tensor; jitter [R,S] (the reference's CPU torch.rand draws, uploaded) for parity,
    or philox_seed for the fast in-kernel RNG, or neither for eval."""
    L.require_cuda(ladder, jitter)
    ladder = L.f32c(ladder)
    S = ladder.shape[0]
    z = _empty((num_rays, S), ladder)
    jitter = L.f32c(jitter)
    if jitter is not None:
        assert jitter.shape == (num_rays, S)
    L.call('srf_stratified_z', L.ptr(ladder), S, num_rays, L.ptr(jitter), int(jitter is None and philox_seed is not None),
           int(philox_seed or 0), L.ptr(z), L.stream_handle())
    return z


def box_march_z(rays_o, rays_d, num_samples, bbox, near, far, step_size, jitter=None):
    """Simple-TensoRF depths without NDC (SimpleTensoRF09.py:388-400): box entry clamped to [near, far] + step_size * (s + jitter[r]).
    bbox: [[min xyz], [max xyz]] host values; jitter: [R] or [R,1] device tensor (one draw per ray) or None."""
    import ctypes
    L.require_cuda(rays_o, rays_d, jitter)
    rays_o, rays_d, jitter = L.f32c(rays_o), L.f32c(rays_d), L.f32c(jitter)
    R = rays_o.shape[0]
    z = _empty((R, int(num_samples)), rays_o)
    box = (ctypes.c_float * 6)(*[float(v) for row in bbox for v in row])
    L.call('srf_box_march_z', L.ptr(rays_o), L.ptr(rays_d), R, int(num_samples), box, float(near), float(far), float(step_size),
           L.ptr(None if jitter is None else jitter.reshape(-1)), L.ptr(z), L.stream_handle())
    return z


def sample_pdf_merge(z_coarse, weights, num_fine, u=None, philox_seed=0, return_indices=False):
    """u: [R,N] tensor, [N] shared row (deterministic linspace) or None (in-kernel Philox).
    Returns z_fine [R,S+N] (+ samples, below, above when return_indices)."""
    L.require_cuda(z_coarse, weights, u)
    z_coarse, weights, u = L.f32c(z_coarse), L.f32c(weights), L.f32c(u)
    R, S = z_coarse.shape
    assert weights.shape == (R, S)
    stride = 0
    if u is not None:
        if u.dim() == 2:
            assert u.shape == (R, num_fine)
            stride = num_fine
        else:
            assert u.shape == (num_fine,)
    z_fine = _empty((R, S + num_fine), z_coarse)
    samples = below = above = None
    if return_indices:
        samples = _empty((R, num_fine), z_coarse)
        below = _empty((R, num_fine), z_coarse, torch.int64)
        above = _empty((R, num_fine), z_coarse, torch.int64)
    L.call('srf_sample_pdf_merge', L.ptr(z_coarse), L.ptr(weights), L.ptr(u), stride, int(philox_seed), R, S, num_fine,
           L.ptr(z_fine), L.ptr(samples), L.ptr(below), L.ptr(above), L.stream_handle())
    if return_indices:
        return z_fine, samples, below, above
    return z_fine


# ------------------------------------------------------------------------------------------------
# compositing (reference: models/SimpleNeRF17.py:486-539, models/SimpleTensoRF09.py:767-819)
# ------------------------------------------------------------------------------------------------
class _Composite(torch.autograd.Function):
    """Outputs: alpha, visibility, weights, rgb_map, acc, depth, depth_var, depth_ndc, depth_var_ndc.
    Differentiable w.r.t. sigma and rgb by the hand-written backward kernel (sample depths carry no gradient in the reference:
    z_samples.detach(), SimpleNeRF17.py:368).  With learnable cameras (SimpleNeRF17.py:817-842; no shipped config) the rays carry
    gradient too: `_composite_ray_gradients` below adds the two places the rays enter compositing directly."""

    @staticmethod
    def forward(ctx, sigma, rgb, z, rays_o, rays_d, rays_d_ndc, ndc, white_bkgd, distance_scale, per_sample):
        L.require_cuda(sigma, rgb, z, rays_o, rays_d, rays_d_ndc)
        needs_grad = any(t is not None and t.requires_grad for t in (sigma, rgb, rays_o, rays_d, rays_d_ndc))
        sigma, rgb, z = L.f32c(sigma), L.f32c(rgb), L.f32c(z)
        rays_o, rays_d, rays_d_ndc = L.f32c(rays_o), L.f32c(rays_d), L.f32c(rays_d_ndc)
        R, S = sigma.shape
        keep = per_sample or needs_grad
        weights = _empty((R, S), sigma)
        alpha = _empty((R, S), sigma) if keep else None
        vis = _empty((R, S), sigma) if keep else None
        rgb_map = _empty((R, 3), sigma) if rgb is not None else None
        acc, depth, depth_var = (_empty((R,), sigma) for _ in range(3))
        depth_ndc = _empty((R,), sigma) if ndc else None
        depth_var_ndc = _empty((R,), sigma) if ndc else None
        L.call('srf_composite_fwd', L.ptr(sigma), L.ptr(rgb), L.ptr(z), L.ptr(rays_o), L.ptr(rays_d), L.ptr(rays_d_ndc),
               R, S, int(ndc), int(white_bkgd), float(distance_scale), L.ptr(alpha), L.ptr(vis), L.ptr(weights),
               L.ptr(rgb_map), L.ptr(acc), L.ptr(depth), L.ptr(depth_var), L.ptr(depth_ndc), L.ptr(depth_var_ndc),
               L.stream_handle(),
               # algorithmic bytes: sigma + z in, weights out (+ rgb in, + alpha & visibility out) per sample
               work=float(R) * S * (12 + (12 if rgb is not None else 0) + (8 if keep else 0)) + float(R) * 68)
        ctx.cfg = (bool(ndc), bool(white_bkgd), float(distance_scale), rgb is not None)
        ctx.save_for_backward(sigma, rgb, z, vis, rays_o, rays_d, rays_d_ndc, acc, depth, depth_ndc)
        outs = (alpha, vis, weights, rgb_map, acc, depth, depth_var, depth_ndc, depth_var_ndc)
        ctx.mark_non_differentiable(*[t for t in (alpha, vis) if t is not None])
        return outs

    @staticmethod
    def backward(ctx, g_alpha, g_vis, g_weights, g_rgb, g_acc, g_depth, g_depth_var, g_depth_ndc, g_depth_var_ndc):
        sigma, rgb, z, vis, rays_o, rays_d, rays_d_ndc, acc, depth, depth_ndc = ctx.saved_tensors
        ndc, white, scale, has_rgb = ctx.cfg
        R, S = sigma.shape
        gs = [L.f32c(g) for g in (g_rgb, g_acc, g_depth, g_depth_ndc, g_depth_var, g_depth_var_ndc, g_weights)]
        g_sigma = _empty((R, S), sigma)
        want_rgb = has_rgb and ctx.needs_input_grad[1] and gs[0] is not None
        g_rgb_s = _empty((R, S, 3), sigma) if want_rgb else None
        L.call('srf_composite_bwd', L.ptr(sigma), L.ptr(rgb if gs[0] is not None else None), L.ptr(z), L.ptr(vis),
               L.ptr(rays_o), L.ptr(rays_d), L.ptr(rays_d_ndc), L.ptr(acc), L.ptr(depth), L.ptr(depth_ndc),
               *[L.ptr(g) for g in gs], R, S, int(ndc), int(white), scale, L.ptr(g_sigma), L.ptr(g_rgb_s),
               L.stream_handle(),
               # sigma, z, visibility (+ rgb) in, g_sigma (+ g_rgb) out (+ a direct weights gradient in) per sample
               work=float(R) * S * (16 + (12 if gs[0] is not None else 0) + (12 if want_rgb else 0) + (4 if gs[6] is not None else 0)))
        if has_rgb and ctx.needs_input_grad[1] and g_rgb_s is None:
            g_rgb_s = torch.zeros((R, S, 3), dtype=torch.float32, device=sigma.device)
        g_z = g_o = g_d = g_dn = None
        if any(ctx.needs_input_grad[2:6]):
            g_o, g_d, g_dn, g_z = _composite_ray_gradients(sigma, z, vis, rays_o, rays_d, rays_d_ndc, acc, ndc, scale, g_sigma,
                                                           g_depth=gs[2], g_depth_var=gs[4], want_z=ctx.needs_input_grad[2])
        return g_sigma, g_rgb_s, g_z, g_o, g_d, g_dn, None, None, None, None


def _ndc_to_world_depth(z_ndc, rays_o, rays_d):
    """SimpleNeRF17.py:541-559 (near plane hard-coded to 1 there as well)."""
    oz, dz = rays_o[..., 2:3], rays_d[..., 2:3]
    tn = -(1 + oz) / dz
    guard = torch.where(z_ndc == 1., 1e-3, 0.)
    return (oz + tn * dz) / dz * (1 / (1 - z_ndc + guard) - 1) + tn


def _composite_ray_gradients(sigma, z, vis, rays_o, rays_d, rays_d_ndc, acc, ndc, scale, g_sigma, g_depth, g_depth_var, want_z=False):
    """Learnable cameras only: gradient of the compositing outputs w.r.t. the rays, sigma / rgb / z held fixed.  The rays enter in two places
    (SimpleNeRF17.py:486-516): (1) delta = dists * |d| — alpha = 1 - exp(-sigma delta) is symmetric in sigma and |d|, so
    dL/d|d| = sum_s g_sigma sigma / |d| with the g_sigma the backward kernel just produced for ALL outputs; (2) with NDC the world depths
    z = convert_depth_from_ndc(z_ndc, rays_o, rays_d) under `depth` / `depth_var`.  [R, S] elementwise torch ops inside backward(): the
    path exists for the test-time pose refinement of Tester07.py:62-111, not for training throughput.
    want_z (world space only: the box-march depths of Simple-TensoRF start at the ray's entry into the box, SimpleTensoRF09.py:388-400, so
    they depend on the pose): also the gradient w.r.t. the sample depths — through the intervals (dL/d dist = g_sigma sigma / dist, the same
    symmetry) and through `depth` / `depth_var` with the weights held fixed.  Returns (g_rays_o, g_rays_d, g_rays_d_ndc, g_z)."""
    if want_z and ndc:
        raise NotImplementedError('depth gradients are only derived for world-space compositing (NDC depths never depend on the cameras)')
    with torch.enable_grad():
        ro = rays_o.detach().requires_grad_(True) if rays_o is not None else None
        rd = rays_d.detach().requires_grad_(True)
        rdn = rays_d_ndc.detach().requires_grad_(True) if (ndc and rays_d_ndc is not None) else None
        norm = torch.norm(rdn if ndc else rd, dim=-1)
        total = ((g_sigma * sigma).sum(-1) / norm.detach() * norm).sum()
        if ndc and (g_depth is not None or g_depth_var is not None):
            dists = torch.cat([z[:, 1:], torch.ones_like(z[:, :1])], -1) - z
            weights = (1. - torch.exp(-sigma * dists * norm.detach()[:, None] * scale)) * vis
            z_world = _ndc_to_world_depth(z, ro, rd)
            depth = torch.sum(weights * z_world, dim=-1) / (acc + 1e-6)
            if g_depth is not None:
                total = total + (g_depth * depth).sum()
            if g_depth_var is not None:
                total = total + (g_depth_var * torch.sum(weights * torch.square(z_world - depth[..., None]), dim=-1)).sum()
        zz = None
        if want_z:
            zz = z.detach().requires_grad_(True)
            dists = torch.cat([zz[:, 1:], torch.full_like(zz[:, :1], 1e10)], -1) - zz
            d0 = dists.detach()
            per_interval = torch.where(d0 != 0, g_sigma * sigma / torch.where(d0 != 0, d0, torch.ones_like(d0)), torch.zeros_like(d0))
            total = total + (per_interval * dists).sum()
            weights = (1. - torch.exp(-sigma * d0 * norm.detach()[:, None] * scale)) * vis
            depth = torch.sum(weights * zz, dim=-1) / (acc + 1e-6)
            if g_depth is not None:
                total = total + (g_depth * depth).sum()
            if g_depth_var is not None:
                total = total + (g_depth_var * torch.sum(weights * torch.square(zz - depth[..., None]), dim=-1)).sum()
        wrt = [t for t in (ro, rd, rdn, zz) if t is not None]
        grads = dict(zip(map(id, wrt), torch.autograd.grad(total, wrt, allow_unused=True)))
    return tuple(None if t is None else grads[id(t)] for t in (ro, rd, rdn, zz))


def composite(sigma, rgb, z, rays_o, rays_d, rays_d_ndc=None, *, ndc, white_bkgd=False, distance_scale=1.0,
              per_sample=True):
    """Returns a dict with the reference's volume_rendering keys.  sigma [R,S], rgb [R,S,3] or None."""
    (alpha, vis, weights, rgb_map, acc, depth, depth_var, depth_ndc, depth_var_ndc) = _Composite.apply(
        sigma, rgb, z, rays_o, rays_d, rays_d_ndc, ndc, white_bkgd, distance_scale, per_sample)
    out = {'acc': acc, 'weights': weights, 'depth': depth, 'depth_var': depth_var}
    if rgb_map is not None:
        out['rgb'] = rgb_map
    if alpha is not None:
        out['alpha'] = alpha
        out['visibility'] = vis
    if ndc:
        out['depth_ndc'] = depth_ndc
        out['depth_var_ndc'] = depth_var_ndc
    return out
