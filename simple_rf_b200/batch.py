"""Host side of the fused batch assembly (csrc/batch.cu), "next" row f2 of SURVEY.md §8f.

Reference: src/data_preprocessors/DataPreprocessor10.py:530-549 (load_nerf_cached_batch) and :568-595
(load_sparse_depth_cached_batch).  No CPU / eager fallback."""
import torch

from . import _lib as L


_ERROR_FLAGS = {}


def _error_flag(device):
    """One int32 in mapped pinned host memory per device: the kernel raises it when an index is out of range, the host reads
    it at the next call without synchronising the stream."""
    f = _ERROR_FLAGS.get(device)
    if f is None:
        f = _ERROR_FLAGS[device] = torch.zeros(1, dtype=torch.int32).pin_memory()
    return f


def check_indices_error(device, synchronize=False):
    """Raises IndexError if a previous assemble_batch() on `device` met an index outside the tables (what the reference's
    fancy indexing does immediately, DataPreprocessor10.py:538)."""
    f = _ERROR_FLAGS.get(device)
    if f is None:
        return
    if synchronize:
        torch.cuda.synchronize(device)
    if int(f[0]) != 0:
        f[0] = 0
        raise IndexError('assemble_batch: a batch index lies outside the cached per-pixel tables')


def assemble_batch(indices, mask_sparse_depth, pixel_table, rgb_table, depth_table=None, error_table=None, points_table=None):
    """indices int64 [B] (flat pixel indices), mask_sparse_depth bool [B] or None; tables [N,3] int32 / [N,3] / [N,1] / [N,1] /
    [N,3] fp32, all CUDA.  Returns dict with pixel_id int32 [B,3], target_rgb [B,3] and, when the sparse-depth tables are
    given, sparse_depth_values [B,1], sparse_depth_errors [B,1], sparse_depth_points3d [B,3] (-1 where a ray kind does not
    carry the field)."""
    L.require_cuda(indices, mask_sparse_depth, pixel_table, rgb_table, depth_table, error_table, points_table)
    assert indices.dtype == torch.int64 and pixel_table.dtype == torch.int32
    indices = indices.contiguous()
    dev = indices.device
    check_indices_error(dev)
    flag = _error_flag(dev)
    B, N = indices.shape[0], pixel_table.shape[0]
    mask = None if mask_sparse_depth is None else mask_sparse_depth.to(torch.uint8).contiguous()
    out = {'pixel_id': torch.empty((B, 3), dtype=torch.int32, device=dev), 'target_rgb': torch.empty((B, 3), dtype=torch.float32, device=dev)}
    sd = depth_table is not None
    if sd:
        out['sparse_depth_values'] = torch.empty((B, 1), dtype=torch.float32, device=dev)
        out['sparse_depth_errors'] = torch.empty((B, 1), dtype=torch.float32, device=dev)
        out['sparse_depth_points3d'] = torch.empty((B, 3), dtype=torch.float32, device=dev)
    f = L.f32c
    tables = [pixel_table.contiguous(), f(rgb_table), f(depth_table), f(error_table), f(points_table)]
    L.call('srf_assemble_batch', L.ptr(indices), L.ptr(mask), B, N, *[L.ptr(t) for t in tables], L.ptr(out['pixel_id']),
           L.ptr(out['target_rgb']), L.ptr(out.get('sparse_depth_values')), L.ptr(out.get('sparse_depth_errors')),
           L.ptr(out.get('sparse_depth_points3d')), flag.data_ptr(), L.stream_handle())
    return out
