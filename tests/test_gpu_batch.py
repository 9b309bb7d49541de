"""GPU parity of the fused batch assembly (f2) through the C ABI: bit-exact against the oracle and the reference golden."""
import pytest
import torch

from oracle import batch as OB

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _run(indices, m_sd, t, with_sd=True):
    from simple_rf_b200 import batch
    d = {k: v.to(DEV) for k, v in t.items()}
    extra = (d['depth'], d['error'], d['points']) if with_sd else ()
    return batch.assemble_batch(indices.to(DEV), None if m_sd is None else m_sd.to(DEV), d['pixel'], d['rgb'], *extra)


def test_against_reference_golden(golden):
    g = golden('batch_assembly')
    out = _run(g['indices'], g['mask_sd'], OB.synthetic_tables())
    for k in ('pixel_id', 'target_rgb', 'sparse_depth_values', 'sparse_depth_errors', 'sparse_depth_points3d'):
        assert torch.equal(out[k].cpu(), g[k]), k


@pytest.mark.parametrize('num_nerf,num_sd', [(0, 0), (1, 0), (0, 5), (2048, 2048), (100000, 37)])
def test_against_oracle(num_nerf, num_sd):
    t = OB.synthetic_tables(num_views=3, h=96, w=128, seed=4)
    indices, m_nerf, m_sd = OB.synthetic_indices(t['pixel'].shape[0], num_nerf, num_sd, seed=num_nerf + num_sd)
    ref = OB.assemble_batch(indices, m_nerf, m_sd, t['pixel'], t['rgb'], t['depth'], t['error'], t['points'])
    out = _run(indices, m_sd, t, with_sd=m_sd is not None)
    assert set(out) == set(ref)
    for k in ref:
        assert out[k].dtype == ref[k].dtype and torch.equal(out[k].cpu(), ref[k]), k


def test_out_of_range_index_is_reported_not_read():
    """The reference's fancy indexing raises IndexError for an index outside the tables (DataPreprocessor10.py:538); the kernel
    fills the row with -1, never reads out of bounds, and the host raises at the next call (or on an explicit check)."""
    from simple_rf_b200 import batch
    t = OB.synthetic_tables(num_views=2, h=16, w=16, seed=1)
    n = t['pixel'].shape[0]
    indices = torch.tensor([0, n - 1, n, -1, 5], dtype=torch.int64)
    out = _run(indices, None, t, with_sd=False)
    torch.cuda.synchronize()
    assert torch.equal(out['pixel_id'].cpu()[[2, 3]], torch.full((2, 3), -1, dtype=torch.int32))
    assert torch.equal(out['pixel_id'].cpu()[[0, 1, 4]], t['pixel'][[0, n - 1, 5]])
    with pytest.raises(IndexError):
        batch.check_indices_error(torch.device('cuda', torch.cuda.current_device()), synchronize=True)
    _run(torch.tensor([1, 2], dtype=torch.int64), None, t, with_sd=False)      # flag was cleared: a clean call passes


def test_frame_output_record_matches_numpy_post_processing():
    """srf_frame_outputs against the reference's CPU post-processing (DataPreprocessor10.py:982-995): uint8 image bit-exact
    (clip, x255 in fp32, round half to even), depth maps bit-exact (negative -> 0), ragged ray counts."""
    import numpy
    from simple_rf_b200 import _lib as L
    lib = L.load()
    for n in (1, 3, 4, 1023, 756 * 1008):
        g = torch.Generator().manual_seed(n)
        rgb = torch.rand(n, 3, generator=g) * 1.4 - 0.2
        rgb[::7] = torch.round(rgb[::7] * 255) / 255 + 0.5 / 255              # exact .5 levels: half-to-even decides
        maps = [torch.randn(n, generator=g) * 3 for _ in range(4)]
        nbytes = int(lib.srf_frame_record_bytes(n, 4))
        rec = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
        dev = [rgb.to(DEV)] + [m.to(DEV) for m in maps]
        L.call('srf_frame_outputs', *[L.ptr(x) for x in dev], n, L.ptr(rec), L.stream_handle())
        buf = rec.cpu().numpy()
        img_b, map_b = (n * 3 + 15) // 16 * 16, (n * 4 + 15) // 16 * 16
        want = numpy.round(numpy.clip(rgb.numpy(), 0, 1) * 255).astype('uint8')
        assert numpy.array_equal(buf[:n * 3].reshape(n, 3), want), n
        for i, m in enumerate(maps):
            got = buf[img_b + i * map_b: img_b + i * map_b + n * 4].view(numpy.float32)
            assert numpy.array_equal(got, numpy.clip(m.numpy(), 0, numpy.inf).astype('float32')), (n, i)
