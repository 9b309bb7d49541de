"""Oracle: the patch-reprojection depth losses that sit right after the render path in every training step
(test infrastructure).

Follows src/loss_functions/AugmentationsDepthLoss11.py:105-190 (`compute_loss_nerf`, main vs augmented depth),
src/loss_functions/CoarseFineConsistencyLoss34.py:89-168 (coarse vs fine depth; the same routine without the
"both reprojections invalid" rule) and src/utils/CommonUtils04.py:227-253 (`reproject`).  Functional torch ops in the
reference's order; the 25 x 3 fancy-index gathers are restated as one gather per patch with the same zero-outside
semantics (the reference pads the images and lets negative indices wrap into the padding).
"""
import torch


def closest_views(poses):
    """AugmentationsDepthLoss11.py:131-135 per view: index of the nearest OTHER camera (2nd smallest distance)."""
    origins = poses[:, :3, 3]
    dist = torch.sqrt(torch.sum(torch.square(origins[:, None, :] - origins[None, :, :]), dim=2))
    return torch.kthvalue(dist, 2, dim=1)[1]


def reproject(points, poses_b, k_first):
    """CommonUtils04.py:227-253: pixel position of world points in the views poses_b [N,4,4] with the FIRST ray's
    intrinsics (the reference hard-codes intrinsics[:1])."""
    origins = poses_b[:, :3, 3]
    rot = poses_b[:, :3, :3]
    d = points - origins
    permuter = torch.eye(3, dtype=points.dtype)
    permuter[1:] *= -1
    pos = (k_first[None] @ permuter[None] @ rot.transpose(1, 2) @ d[..., None]).squeeze(-1)
    return pos[:, :2] / pos[:, 2:]


def _patches(images, view, x, y, hpx, hpy):
    """[N, py, px, C] patches centred on (x, y) of images[view]; zeros outside the frame."""
    v, h, w, c = images.shape
    oy = torch.arange(-hpy, hpy + 1)
    ox = torch.arange(-hpx, hpx + 1)
    yy = y[:, None, None] + oy[None, :, None]
    xx = x[:, None, None] + ox[None, None, :]
    inside = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
    val = images[view[:, None, None], yy.clamp(0, h - 1), xx.clamp(0, w - 1)]
    return val * inside[..., None].to(images.dtype)


def patch_reprojection_masks(rays_o, rays_d, depth1, depth2, pixel_id, poses, k_first, images, patch_size=(5, 5),
                             rmse_threshold=0.1, both_invalid_rule=True):
    """mask1 / mask2 [N] bool (mask1: model 1 is the more accurate one) and the two patch RMSEs, for the N image rays.
    pixel_id [N,3] = (view, x, y) integer."""
    px, py = patch_size
    hpx, hpy = px // 2, py // 2
    v, h, w, _ = images.shape
    pixel_id = pixel_id.long()
    view_a = pixel_id[:, 0]
    view_b = closest_views(poses)[view_a]
    poses_b = poses[view_b]
    p1 = rays_o + rays_d * depth1[:, None]
    p2 = rays_o + rays_d * depth2[:, None]
    pos1 = reproject(p1.detach(), poses_b, k_first).round().long()
    pos2 = reproject(p2.detach(), poses_b, k_first).round().long()
    xa, ya = pixel_id[:, 1], pixel_id[:, 2]

    def valid(x, y):
        return (x >= hpx) & (x < w - hpx) & (y >= hpy) & (y < h - hpy)
    va, v1, v2 = valid(xa, ya), valid(pos1[:, 0], pos1[:, 1]), valid(pos2[:, 0], pos2[:, 1])
    pa = _patches(images, view_a, xa, ya, hpx, hpy)
    p1b = _patches(images, view_b, pos1[:, 0].clip(0, w - 1), pos1[:, 1].clip(0, h - 1), hpx, hpy)
    p2b = _patches(images, view_b, pos2[:, 0].clip(0, w - 1), pos2[:, 1].clip(0, h - 1), hpx, hpy)
    rmse1 = torch.sqrt(torch.mean(torch.square(pa - p1b), dim=(1, 2, 3)))
    rmse2 = torch.sqrt(torch.mean(torch.square(pa - p2b), dim=(1, 2, 3)))
    mask1 = ((rmse1 < rmse2) | ~v2) & (rmse1 < rmse_threshold) & v1 & va
    mask2 = ((rmse2 < rmse1) | ~v1) & (rmse2 < rmse_threshold) & v2 & va
    if both_invalid_rule:                                     # AugmentationsDepthLoss11.py:178-182
        both_invalid = ~(v1 | v2)
        mask1 = mask1 | (both_invalid & (depth1 > depth2))
        mask2 = mask2 | (both_invalid & (depth2 > depth1))
    return mask1, mask2, rmse1, rmse2


def masked_depth_loss(depth1, depth2, mask1, mask2):
    """AugmentationsDepthLoss11.py:184-189 + :208-224, INCLUDING its aliasing quirk: `compute_depth_mse` zeroes its
    arguments in place, and the first call receives `depth2.detach()` (same storage as depth2), so by the time the second
    call runs both depths are already zero outside mask2.  Net effect: the first term pulls depth1 towards depth2 where
    model 2 is the more accurate; the second term only survives where mask1 AND mask2 hold (never, by construction of
    the masks), so depth2 receives no gradient.  Means are over all N rays."""
    m1, m2 = mask1.to(depth1.dtype), mask2.to(depth1.dtype)
    map1 = torch.square(depth1 * m2 - depth2.detach() * m2)
    map2 = torch.square(depth2 * (m1 * m2) - depth1.detach() * (m1 * m2))
    zero = torch.zeros((), dtype=depth1.dtype)
    loss = (map1.mean() if depth1.numel() > 0 else zero) + (map2.mean() if depth1.numel() > 0 else zero)
    return loss, map1, map2


def tv_loss(planes, iter_weight):
    """src/loss_functions/TotalVariationLoss04.py:97-116 (compute_tv_loss) over a list of [1,C,H,W] planes."""
    total = 0
    for c in planes:
        dh = torch.pow(c[:, :, 1:, :] - c[:, :, :-1, :], 2)
        dw = torch.pow(c[:, :, :, 1:] - c[:, :, :, :-1], 2)
        total = total + 2 * (dh.sum() / max(dh.numel(), 1) + dw.sum() / max(dw.numel(), 1)) * iter_weight
    return total
