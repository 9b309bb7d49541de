"""Oracle: stratified coarse samples and hierarchical inverse-CDF resampling (test infrastructure).

Follows src/models/SimpleNeRF17.py:330-358 (get_z_vals_coarse), :360-371 (get_z_vals_fine),
:385-417 (sample_pdf) — identical code at src/models/SimpleTensoRF09.py:353-381, :403-460 —
and the TensoRF ray/box marching variant at SimpleTensoRF09.py:388-400.
"""
import torch


def coarse_depths(num_samples, near, far, lindisp=False):
    """SimpleNeRF17.py:341-345: the shared, un-jittered [S] depth ladder."""
    t = torch.linspace(0., 1., steps=num_samples)
    if not lindisp:
        return near * (1. - t) + far * t
    return 1. / (1. / near * (1. - t) + 1. / far * t)


def stratified_depths(ladder, num_rays, jitter=None):
    """SimpleNeRF17.py:347-357.  `jitter` is the [R,S] uniform tensor the reference draws with
    torch.rand on the CPU generator (:355); None means eval / perturb=False."""
    z = ladder.expand([num_rays, ladder.shape[0]])
    if jitter is None:
        return z
    mid = .5 * (z[..., 1:] + z[..., :-1])
    hi = torch.cat([mid, z[..., -1:]], -1)
    lo = torch.cat([z[..., :1], mid], -1)
    return lo + (hi - lo) * jitter


def box_march_depths(rays_o, rays_d, bbox, near, far, step_size, num_samples, jitter=None):
    """SimpleTensoRF09.py:388-400 (non-NDC).  `jitter` [R,1] is torch.rand_like(rng[:, [0]]) (:398)."""
    safe_d = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
    ra = (bbox[1] - rays_o) / safe_d
    rb = (bbox[0] - rays_o) / safe_d
    t_min = torch.minimum(ra, rb).amax(-1).clamp(min=near, max=far)
    rng = torch.arange(num_samples)[None].float()
    if jitter is not None:
        rng = rng.repeat(rays_d.shape[-2], 1) + jitter
    return t_min[..., None] + step_size * rng


def inverse_cdf(bins, weights, u):
    """SimpleNeRF17.py:385-417 with the uniform draws `u` [R,N] passed in
    (linspace(0,1,N) expanded when det, torch.rand on the CPU generator otherwise, :393-397).
    Returns (samples [R,N], below int64 [R,N], above int64 [R,N], cdf [R,len(bins)])."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b = torch.gather(cdf, 1, below)
    cdf_a = torch.gather(cdf, 1, above)
    bin_b = torch.gather(bins, 1, below)
    bin_a = torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    samples = bin_b + t * (bin_a - bin_b)
    return samples, below, above, cdf


def fine_depths(z_coarse, weights_coarse, u):
    """SimpleNeRF17.py:360-371: midpoints as bins, interior weights, sort(cat(coarse, new)).
    Returns (z_fine [R,S+N], samples, below, above)."""
    mids = .5 * (z_coarse[..., 1:] + z_coarse[..., :-1])
    samples, below, above, _ = inverse_cdf(mids, weights_coarse[..., 1:-1], u)
    samples = samples.detach()                                   # SimpleNeRF17.py:368
    z_fine, _ = torch.sort(torch.cat([z_coarse, samples], -1), -1)
    return z_fine, samples, below, above


def det_u(num_rays, num_fine):
    """SimpleNeRF17.py:394-395."""
    return torch.linspace(0., 1., steps=num_fine).expand([num_rays, num_fine]).contiguous()
