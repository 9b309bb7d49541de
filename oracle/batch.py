"""TEST INFRASTRUCTURE (CPU oracle): restatement of the reference's per-iteration batch assembly,
src/data_preprocessors/DataPreprocessor10.py:530-549 (load_nerf_cached_batch) and :568-595 (load_sparse_depth_cached_batch),
pinned bit-exact against those two methods of the unmodified class in tests/test_batch_cpu.py (golden
tests/golden/batch_assembly.npz, written by oracle/generate_golden.py::golden_batch_assembly)."""
import torch


def assemble_batch(indices, mask_nerf, mask_sd, pixel_table, rgb_table, depth_table=None, error_table=None, points_table=None):
    n = indices.shape[0]
    pixel_id = -1 * torch.ones((n, 3), dtype=torch.int32)                      # :536-537
    target_rgb = -1 * torch.ones((n, 3))
    idx_nerf = indices[mask_nerf]
    pixel_id[mask_nerf] = pixel_table[idx_nerf]                                # :540-541
    target_rgb[mask_nerf] = rgb_table[idx_nerf]
    out = {'pixel_id': pixel_id, 'target_rgb': target_rgb}
    if mask_sd is not None:
        idx_sd = indices[mask_sd]
        pixel_id[mask_sd] = pixel_table[idx_sd]                                # :581
        depths, errors, points = -1 * torch.ones((n, 1)), -1 * torch.ones((n, 1)), -1 * torch.ones((n, 3))   # :584-586
        depths[mask_sd] = depth_table[idx_sd]                                  # :588-590
        errors[mask_sd] = error_table[idx_sd]
        points[mask_sd] = points_table[idx_sd]
        out.update(sparse_depth_values=depths, sparse_depth_errors=errors, sparse_depth_points3d=points)
    return out


def synthetic_tables(num_views=3, h=24, w=32, seed=0):
    """Per-pixel tables of the shapes DataPreprocessor10.preprocess_* caches (:343-410): flat index = view * h * w + y * w + x."""
    g = torch.Generator().manual_seed(seed)
    v, y, x = torch.meshgrid(torch.arange(num_views), torch.arange(h), torch.arange(w), indexing='ij')
    pixel = torch.stack([v, x, y], -1).reshape(-1, 3).int()
    n = pixel.shape[0]
    return {'pixel': pixel, 'rgb': torch.rand(n, 3, generator=g), 'depth': torch.rand(n, 1, generator=g) * 5 + 0.5,
            'error': torch.rand(n, 1, generator=g), 'points': torch.randn(n, 3, generator=g)}


def synthetic_indices(n_pixels, num_nerf, num_sd, seed):
    g = torch.Generator().manual_seed(seed)
    indices = torch.cat([torch.randint(0, n_pixels, (num_nerf,), generator=g), torch.randint(0, n_pixels, (num_sd,), generator=g)])
    ids = torch.cat([torch.ones(num_nerf), 2 * torch.ones(num_sd)])
    return indices, ids == 1, (ids == 2) if num_sd else None
