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
