"""Oracle: ray generation, NDC warp, view directions (test infrastructure, see oracle/__init__.py).

Follows src/utils/CommonUtils04.py:73-95 (get_rays_tr), :120-138 (get_ndc_rays_tr),
:147-149 (get_view_dirs_tr) and the TensoRF x-flip at src/models/SimpleTensoRF09.py:205-207.
"""
import torch


def camera_rays(pixel_id, intrinsics, c2w, *, half_pixel, flip_x):
    """pixel_id int[R,3] = (image, x, y); intrinsics [F,3,3]; c2w [F,4,4] (per-view tables).

    CommonUtils04.py:80-95: p = (x, y, 1) (+0.5 when `mip_nerf_used`), dirs = K^-1 p,
    dirs[1:] *= -1, rays_d = R dirs, rays_o = t.  `flip_x` is SimpleTensoRF09.py:206-207.
    Returns (rays_o [R,3], rays_d [R,3]).
    """
    img = pixel_id[:, 0].long()
    xy = pixel_id[:, 1:].float()
    if half_pixel:
        xy = xy + 0.5
    homog = torch.cat([xy, torch.ones_like(xy[:, :1])], dim=1)              # (R,3)
    k_inv = torch.linalg.inv(intrinsics.float())[img]                       # per-view inverse, gathered
    dirs = (k_inv @ homog[:, :, None])[:, :, 0]
    dirs = dirs * dirs.new_tensor([1.0, -1.0, -1.0])
    rot = c2w[img, :3, :3].float()
    rays_d = (dirs[:, None, :] * rot).sum(-1)
    rays_o = c2w[img, :3, 3].float().clone()
    if flip_x:
        sign = rays_o.new_tensor([-1.0, 1.0, 1.0])
        rays_o = rays_o * sign
        rays_d = rays_d * sign
    return rays_o, rays_d


def ndc_rays(rays_o, rays_d, height, width, fx, fy, near):
    """CommonUtils04.py:120-138.  fx, fy are per-ray tensors [R] (K[0,0], K[1,1])."""
    t = -(near + rays_o[:, 2]) / rays_d[:, 2]
    o = rays_o + t[:, None] * rays_d
    sx = -1. / (width / (2. * fx))
    sy = -1. / (height / (2. * fy))
    o0 = sx * o[:, 0] / o[:, 2]
    o1 = sy * o[:, 1] / o[:, 2]
    o2 = 1. + 2. * near / o[:, 2]
    d0 = sx * (rays_d[:, 0] / rays_d[:, 2] - o[:, 0] / o[:, 2])
    d1 = sy * (rays_d[:, 1] / rays_d[:, 2] - o[:, 1] / o[:, 2])
    d2 = -2. * near / o[:, 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def view_dirs(rays_d):
    """CommonUtils04.py:147-149."""
    return rays_d / torch.linalg.norm(rays_d, ord=2, dim=-1, keepdim=True)


def depth_from_ndc(z_ndc, rays_o, rays_d):
    """CommonUtils04.py:208-224 == SimpleNeRF17.py:542-558 (near hard-coded to 1, +1e-3 where z == 1)."""
    oz = rays_o[..., 2:3]
    dz = rays_d[..., 2:3]
    tn = -(1 + oz) / dz
    eps = torch.where(z_ndc == 1., 1e-3, 0.)
    return (oz + tn * dz) / dz * (1 / (1 - z_ndc + eps) - 1) + tn


def pose_correction(initial_extrinsics, r, t):
    """ExtrinsicsLearner.forward for every view (SimpleNeRF17.py:831-842, :868-912): the learnable pose correction
    inv([Exp(r) | t; 0 0 0 1]) applied from the right to the initial camera-to-world matrices.  r (axis-angle), t: [V, 3];
    Exp is Rodrigues' formula with the reference's 1e-15 guard on |r|.  Differentiable w.r.t. r and t."""
    zero = torch.zeros((r.shape[0], 1), dtype=torch.float32)
    k0 = torch.cat([zero, -r[:, 2:3], r[:, 1:2]], dim=1)
    k1 = torch.cat([r[:, 2:3], zero, -r[:, 0:1]], dim=1)
    k2 = torch.cat([-r[:, 1:2], r[:, 0:1], zero], dim=1)
    skew = torch.stack([k0, k1, k2], dim=2)
    n = r.norm(dim=1) + 1e-15
    rot = torch.eye(3)[None] + (torch.sin(n) / n)[:, None, None] * skew + ((1 - torch.cos(n)) / n ** 2)[:, None, None] * (skew @ skew)
    top = torch.cat([rot, t.unsqueeze(2)], dim=2)
    bottom = torch.zeros_like(top[:, 0:1])
    bottom[:, 0, 3] = 1.0
    return initial_extrinsics.float() @ torch.linalg.inv(torch.cat([top, bottom], dim=1))
