"""Host side of the fused patch-reprojection masks (csrc/losses.cu) and the masked depth loss built on them.

Reference: src/loss_functions/AugmentationsDepthLoss11.py:105-190, src/loss_functions/CoarseFineConsistencyLoss34.py:89-168,
src/utils/CommonUtils04.py:227-253.  No CPU / eager fallback.
"""
import torch

from .. import _lib as L


def closest_views(poses):
    """Nearest OTHER camera of every view (AugmentationsDepthLoss11.py:131-135: 2nd smallest distance, `kthvalue`); the
    reference evaluates this per ray, the rows of one view are identical."""
    origins = poses[:, :3, 3]
    dist = torch.sqrt(torch.sum(torch.square(origins[:, None, :] - origins[None, :, :]), dim=2))
    return torch.kthvalue(dist, 2, dim=1)[1].to(torch.int32)


def patch_reprojection_masks(rays_o, rays_d, depth1, depth2, pixel_id, poses, intrinsics_first, images, patch_size,
                             rmse_threshold, both_invalid_rule, return_rmse=False):
    """mask1 / mask2 [N] bool for N image rays (mask1: model 1 is the more accurate).  pixel_id [N,3] (view, x, y);
    poses [V,4,4]; intrinsics_first [3,3] (device); images [V,H,W,3] fp32."""
    L.require_cuda(rays_o, rays_d, depth1, depth2, pixel_id, poses, intrinsics_first, images)
    n = depth1.shape[0]
    dev = depth1.device
    mask1 = torch.empty((n,), dtype=torch.uint8, device=dev)
    mask2 = torch.empty((n,), dtype=torch.uint8, device=dev)
    r1 = torch.empty((n,), dtype=torch.float32, device=dev) if return_rmse else None
    r2 = torch.empty((n,), dtype=torch.float32, device=dev) if return_rmse else None
    v, h, w, c = images.shape
    assert c == 3, 'rgb images expected'
    pid = pixel_id.to(torch.int32).contiguous()
    closest = closest_views(poses.detach())
    L.call('srf_patch_reprojection_masks', L.ptr(L.f32c(rays_o.detach())), L.ptr(L.f32c(rays_d.detach())), L.ptr(L.f32c(depth1.detach())),
           L.ptr(L.f32c(depth2.detach())), L.ptr(pid), n, L.ptr(closest), L.ptr(L.f32c(poses.detach())),
           L.ptr(L.f32c(intrinsics_first.detach())), L.ptr(L.f32c(images)), v, h, w, int(patch_size[0]), int(patch_size[1]),
           float(rmse_threshold), int(both_invalid_rule), L.ptr(mask1), L.ptr(mask2), L.ptr(r1), L.ptr(r2), L.stream_handle())
    out = (mask1.view(torch.bool), mask2.view(torch.bool))
    return out + (r1, r2) if return_rmse else out


def masked_depth_loss(depth1, depth2, mask1, mask2):
    """The reference's two `compute_depth_mse` calls (AugmentationsDepthLoss11.py:184-189) INCLUDING their aliasing
    quirk: the first call zeroes `depth2.detach()` in place, so the second term only survives where mask1 AND mask2
    hold - never, by construction of the masks - and depth2 receives no gradient (pinned against the reference by
    oracle/generate_golden.py).  Returns (loss, map1, map2)."""
    m2 = mask2.to(depth1.dtype)
    m12 = (mask1 & mask2).to(depth1.dtype)
    map1 = torch.square((depth1 - depth2.detach()) * m2)
    map2 = torch.square((depth2 - depth1.detach()) * m12)
    if depth1.numel() == 0:
        zero = torch.zeros((), dtype=depth1.dtype, device=depth1.device)
        return zero, map1, map2
    return map1.mean() + map2.mean(), map1, map2


def batch_rows(input_dict, kind):
    """Row indices of the image rays ('nerf') / sparse-depth rays of the batch when the preprocessor supplied them
    (`srf_rows` = {'nerf': int64 tensor, 'sparse_depth': int64 tensor}, DataPreprocessor91), else None."""
    rows = input_dict.get('srf_rows')
    if rows is None:
        return None
    return rows.get(kind)


def consistency_loss_nerf(depth1, depth2, indices_mask_nerf, rays_o, rays_d, poses, images, pixel_ids, intrinsics, patch_size,
                          rmse_threshold, both_invalid_rule, rows_nerf=None):
    """`compute_loss_nerf` of both reference losses: image rays only (indices_mask_nerf), masks from the fused kernel.
    rows_nerf: optional int64 row indices equivalent to the mask (DataPreprocessor91 knows them on the host) — boolean-mask
    indexing has to read the mask's population count back, i.e. synchronises the stream once per indexed tensor."""
    if rows_nerf is not None:
        sel = lambda t: t.index_select(0, rows_nerf)
    else:
        sel = lambda t: t[indices_mask_nerf]
    d1, d2 = sel(depth1), sel(depth2)
    mask1, mask2 = patch_reprojection_masks(sel(rays_o), sel(rays_d), d1, d2, sel(pixel_ids), poses, intrinsics[0], images, patch_size,
                                            rmse_threshold, both_invalid_rule)
    return masked_depth_loss(d1, d2, mask1, mask2)
