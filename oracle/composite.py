"""Oracle: volume-rendering compositing, forward and closed-form backward (test infrastructure).

Forward follows src/models/SimpleNeRF17.py:486-539 (volume_rendering), which is the same
arithmetic as src/models/SimpleTensoRF09.py:767-819 (get_volume_rendering_weights +
volume_render) once `distance_scale` (:787) is folded in.  The backward is the closed form
SURVEY.md §8a row VIII derives; tests check it against autograd through `composite()`.
"""
import torch

from .rays import depth_from_ndc


def composite(sigma, rgb, z, rays_o, rays_d, rays_d_ndc=None, *, ndc, distance_scale=1.0,
              white_bkgd=False):
    """sigma [R,S], rgb [R,S,3] or None, z [R,S] (NDC depths when `ndc`).
    Returns dict with alpha, visibility, weights [R,S]; acc, depth, depth_var (+ depth_ndc,
    depth_var_ndc when ndc) [R]; rgb [R,3] when rgb is given."""
    if not ndc:
        inf_depth = 1e10
        scale_dir = rays_d
    else:
        inf_depth = 1
        scale_dir = rays_d_ndc
    last = torch.tensor([inf_depth], dtype=z.dtype).expand(z[..., :1].shape)
    z1 = torch.cat([z, last], -1)
    delta = (z1[..., 1:] - z1[..., :-1]) * torch.norm(scale_dir[..., None, :], dim=-1)

    if distance_scale == 1.0:
        alpha = 1. - torch.exp(-sigma * delta)                               # SimpleNeRF17.py:502
    else:
        alpha = 1. - torch.exp(-sigma * delta * distance_scale)              # SimpleTensoRF09.py:787
    ones = torch.ones((alpha.shape[0], 1), dtype=alpha.dtype)
    visibility = torch.cumprod(torch.cat([ones, 1. - alpha + 1e-10], -1), -1)[:, :-1]
    weights = alpha * visibility
    acc = torch.sum(weights, dim=-1)
    out = {'alpha': alpha, 'visibility': visibility, 'weights': weights, 'acc': acc}
    if not ndc:
        depth = torch.sum(weights * z, dim=-1) / (acc + 1e-6)
        out['depth'] = depth
        out['depth_var'] = torch.sum(weights * torch.square(z - depth[..., None]), dim=-1)
    else:
        depth_ndc = torch.sum(weights * z, dim=-1) / (acc + 1e-6)
        out['depth_ndc'] = depth_ndc
        out['depth_var_ndc'] = torch.sum(weights * torch.square(z - depth_ndc[..., None]), dim=-1)
        zw = depth_from_ndc(z, rays_o, rays_d)
        depth = torch.sum(weights * zw, dim=-1) / (acc + 1e-6)
        out['depth'] = depth
        out['depth_var'] = torch.sum(weights * torch.square(zw - depth[..., None]), dim=-1)
    if rgb is not None:
        rgb_map = torch.sum(weights[..., None] * rgb, dim=-2)
        if white_bkgd:
            rgb_map = rgb_map + (1. - acc[..., None])
        out['rgb'] = rgb_map
    return out


def composite_backward(sigma, rgb, z, rays_o, rays_d, rays_d_ndc, *, ndc, distance_scale=1.0,
                       white_bkgd=False, g_rgb=None, g_acc=None, g_depth=None, g_depth_ndc=None,
                       g_depth_var=None, g_depth_var_ndc=None, g_weights=None):
    """Closed-form gradients (g_sigma [R,S], g_rgb_samples [R,S,3]) of the scalar
    sum(g_rgb*rgb) + sum(g_acc*acc) + ... with respect to sigma and the per-sample colours.
    Computed in the dtype of the inputs (tests use float64 to pin the formula)."""
    dt = sigma.dtype
    R, S = sigma.shape
    fwd = composite(sigma, rgb, z, rays_o, rays_d, rays_d_ndc, ndc=ndc,
                    distance_scale=distance_scale, white_bkgd=white_bkgd)
    alpha, T, w, acc = fwd['alpha'], fwd['visibility'], fwd['weights'], fwd['acc']
    scale_dir = rays_d_ndc if ndc else rays_d
    inf_depth = 1 if ndc else 1e10
    z1 = torch.cat([z, torch.full((R, 1), inf_depth, dtype=dt)], -1)
    delta = (z1[..., 1:] - z1[..., :-1]) * torch.norm(scale_dir, dim=-1, keepdim=True)
    q = 1. - alpha + 1e-10
    A = (acc + 1e-6)[:, None]

    zero = torch.zeros(R, dtype=dt)
    g_acc_t = zero.clone() if g_acc is None else g_acc.clone()
    gw = torch.zeros(R, S, dtype=dt) if g_weights is None else g_weights.clone()
    if g_rgb is not None:
        gw = gw + (rgb * g_rgb[:, None, :]).sum(-1)
        if white_bkgd:
            g_acc_t = g_acc_t - g_rgb.sum(-1)
    gw = gw + g_acc_t[:, None]

    def depth_terms(zz, depth, g_d, g_v):
        out = torch.zeros(R, S, dtype=dt)
        dz = zz - depth[:, None]
        if g_d is not None:
            out = out + g_d[:, None] * dz / A
        if g_v is not None:
            n = (w * zz).sum(-1)
            out = out + g_v[:, None] * (dz * dz - 2 * (n - depth * acc)[:, None] * dz / A)
        return out

    if ndc:
        gw = gw + depth_terms(z, fwd['depth_ndc'], g_depth_ndc, g_depth_var_ndc)
        gw = gw + depth_terms(depth_from_ndc(z, rays_o, rays_d), fwd['depth'], g_depth, g_depth_var)
    else:
        gw = gw + depth_terms(z, fwd['depth'], g_depth, g_depth_var)

    gww = gw * w
    suffix = torch.flip(torch.cumsum(torch.flip(gww, [-1]), -1), [-1]) - gww     # sum_{k>i} g_w[k] w_k
    g_alpha = gw * T - suffix / q
    g_sigma = g_alpha * delta * distance_scale * (1. - alpha)
    g_rgb_s = None if g_rgb is None else w[..., None] * g_rgb[:, None, :]
    return g_sigma, g_rgb_s
