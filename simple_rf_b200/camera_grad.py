"""Learnable cameras: gradients through ray generation (SimpleNeRF17.py:817-842 `ExtrinsicsLearner`, used by the test-time pose
refinement of Tester07.py:62-111 and by `learn_camera_rotation / learn_camera_translation` in training; no shipped config enables them).

The ray VALUES still come from the raygen kernel (`srf_raygen`, bit-identical to the frozen-camera path).  The gradient reaches the pose
correction r, t through three places, each handled where it arises:

  * the sample points pts = o + z d and the view directions that feed the MLP: `srf_nerf_mlp_input_grad` (csrc/nerf_mlp_input_grad.cu)
    from the dZ images the data-gradient chain leaves in HBM, reduced over the samples of a ray (`nerf_program.mlp_input_backward`);
  * |d| under delta = dists |d| and the NDC -> world depth conversion inside compositing: `ops._composite_ray_gradients`;
  * rays -> per-view camera matrices: `_RaysFromCameras.backward` below re-evaluates CommonUtils04.py:73-149 in differentiable torch ops on the
    [R, 3] ray arrays and lets autograd reduce the five ray gradients to the [V, 4, 4] view matrices.  From there on the learner's own torch
    graph (Exp, inverse, product with the initial pose: V-sized) carries it to r and t.
"""
import torch

from . import ops


def rays_from_cameras(c2w_all, pixel_id, intrinsics, height, width, near, *, half_pixel, flip_x, ndc, viewdirs_from_ndc):
    """Differentiable restatement of get_rays_tr / get_ndc_rays_tr / get_view_dirs_tr (CommonUtils04.py:73-149) with the options of
    `srf_raygen`.  c2w_all [V, 4, 4] (may require grad), intrinsics [V, 3, 3], pixel_id [R, 3] (view, x, y)."""
    img = pixel_id[:, 0].long()
    x, y = pixel_id[:, 1].float(), pixel_id[:, 2].float()
    if half_pixel:
        x, y = x + 0.5, y + 0.5
    k = intrinsics.detach().float()
    homo = torch.stack([x, y, torch.ones_like(x)], dim=1)
    dirs = (torch.linalg.inv(k)[img] @ homo[:, :, None])[:, :, 0] * homo.new_tensor([1., -1., -1.])
    e = c2w_all[img]
    rays_d = torch.sum(dirs[:, None, :] * e[:, :3, :3], dim=-1)
    rays_o = e[:, :3, 3]
    if flip_x:
        sign = homo.new_tensor([-1., 1., 1.])
        rays_o, rays_d = rays_o * sign, rays_d * sign
    o_ndc = d_ndc = None
    source = rays_d
    if ndc:
        fx, fy = k[img, 0, 0], k[img, 1, 1]
        sx, sy = -1. / (width / (2. * fx)), -1. / (height / (2. * fy))
        t = -(near + rays_o[:, 2]) / rays_d[:, 2]
        o = rays_o + t[:, None] * rays_d
        o_ndc = torch.stack([sx * o[:, 0] / o[:, 2], sy * o[:, 1] / o[:, 2], 1. + 2. * near / o[:, 2]], -1)
        d_ndc = torch.stack([sx * (rays_d[:, 0] / rays_d[:, 2] - o[:, 0] / o[:, 2]),
                             sy * (rays_d[:, 1] / rays_d[:, 2] - o[:, 1] / o[:, 2]), -2. * near / o[:, 2]], -1)
        if viewdirs_from_ndc:
            source = d_ndc
    view_dirs = source / torch.linalg.norm(source, ord=2, dim=-1, keepdim=True)
    return rays_o, rays_d, o_ndc, d_ndc, view_dirs


class _RaysFromCameras(torch.autograd.Function):
    """forward: the raygen kernel on the current view matrices; backward: the five ray gradients -> view matrices."""

    @staticmethod
    def forward(ctx, c2w_all, pixel_id, intrinsics, k_inv, focal, height, width, near, flags):
        tables = (k_inv, c2w_all.detach().reshape(-1, 16).contiguous(), focal)
        rays_o, rays_d, o_ndc, d_ndc, view_dirs = ops.raygen(pixel_id, tables, height, width, near, **flags)
        ctx.save_for_backward(c2w_all, pixel_id, intrinsics)
        ctx.cfg = (height, width, near, flags)
        if flags['ndc']:
            return rays_o, rays_d, o_ndc, d_ndc, view_dirs
        return rays_o, rays_d, view_dirs

    @staticmethod
    def backward(ctx, *grads):
        c2w_all, pixel_id, intrinsics = ctx.saved_tensors
        height, width, near, flags = ctx.cfg
        with torch.enable_grad():
            cams = c2w_all.detach().requires_grad_(True)
            outs = rays_from_cameras(cams, pixel_id, intrinsics, height, width, near, **flags)
            outs = [o for o in outs if o is not None]
            pairs = [(o, g) for o, g in zip(outs, grads) if g is not None]
            if not pairs:
                return (None,) * 9
            g_cams, = torch.autograd.grad([o for o, _ in pairs], [cams], [g for _, g in pairs])
        return (g_cams,) + (None,) * 8


def rays_with_camera_gradient(c2w_all, pixel_id, intrinsics, k_inv, focal, height, width, near, *, half_pixel, flip_x, ndc,
                              viewdirs_from_ndc):
    """-> rays_o, rays_d, o_ndc, d_ndc, view_dirs like `ops.raygen`, attached to the autograd graph of `c2w_all` [V, 4, 4]."""
    flags = dict(half_pixel=half_pixel, flip_x=flip_x, ndc=ndc, viewdirs_from_ndc=viewdirs_from_ndc)
    out = _RaysFromCameras.apply(c2w_all, pixel_id, intrinsics, k_inv, focal, int(height), int(width), float(near), flags)
    if ndc:
        return out
    return out[0], out[1], None, None, out[2]


def box_entry_depth(rays_o, rays_d, bbox, near, far):
    """The differentiable part of the world-space box march (SimpleTensoRF09.py:390-394): depth at which the ray enters the tensor's
    bounding box, clamped to [near, far].  bbox: [[min xyz], [max xyz]] host values.  `z = z_kernel + (t - t.detach())[:, None]` ties the
    depths `srf_box_march_z` produced to the rays without changing their values (the steps behind the entry do not depend on the ray)."""
    box = torch.as_tensor(bbox, dtype=rays_o.dtype, device=rays_o.device)
    safe_d = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
    rate_a, rate_b = (box[1] - rays_o) / safe_d, (box[0] - rays_o) / safe_d
    return torch.minimum(rate_a, rate_b).amax(-1).clamp(min=near, max=far)
