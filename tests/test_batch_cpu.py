"""CPU suite for the "next" row f2 (device-side batch assembly): the oracle against the committed outputs of the reference's
own load_nerf_cached_batch / load_sparse_depth_cached_batch, and the drop-in preprocessor class resolving through the
reference's unmodified factory."""
import pytest
import torch

from oracle import batch as OB


def test_oracle_matches_reference_golden(golden):
    g = golden('batch_assembly')
    t = OB.synthetic_tables()
    out = OB.assemble_batch(g['indices'], g['mask_nerf'], g['mask_sd'], t['pixel'], t['rgb'], t['depth'], t['error'], t['points'])
    for k, v in out.items():
        assert torch.equal(v, g[k]), k
    assert (out['target_rgb'][g['mask_sd']] == -1).all() and (out['sparse_depth_values'][g['mask_nerf']] == -1).all()
    assert (out['pixel_id'] >= 0).all()


def test_oracle_without_sparse_depth():
    t = OB.synthetic_tables(seed=2)
    indices, m_nerf, m_sd = OB.synthetic_indices(t['pixel'].shape[0], 64, 0, seed=5)
    out = OB.assemble_batch(indices, m_nerf, m_sd, t['pixel'], t['rgb'])
    assert set(out) == {'pixel_id', 'target_rgb'} and torch.equal(out['target_rgb'], t['rgb'][indices])


@pytest.mark.needs_reference
def test_dropin_preprocessor_resolves_through_reference_factory():
    import inspect
    from oracle import reference_harness as H
    from simple_rf_b200 import dropin
    H.import_reference()
    dropin.install()
    import importlib
    module = importlib.import_module('data_preprocessors.DataPreprocessor91')      # what DataPreprocessorFactory01.py:18 does
    classes = dict(inspect.getmembers(module, inspect.isclass))
    cls = classes['DataPreprocessor91'[:-2]]
    from data_preprocessors.DataPreprocessor10 import DataPreprocessor as Ref
    assert issubclass(cls, Ref) and cls.__module__.startswith('simple_rf_b200.data_preprocessors')
    assert cls.load_nerf_cached_batch is not Ref.load_nerf_cached_batch
    # single process: the override delegates to the reference (RNG / shuffling order stays the reference's); the multi-rank
    # slicing is covered by tests/test_parallel_cpu.py::test_preprocessor_rank_sharding_gloo
    assert cls.generate_indices is Ref.generate_indices


def test_assemble_batch_refuses_cpu():
    from simple_rf_b200 import batch
    t = OB.synthetic_tables()
    indices, m_nerf, m_sd = OB.synthetic_indices(t['pixel'].shape[0], 8, 8, seed=1)
    with pytest.raises(RuntimeError):
        batch.assemble_batch(indices, m_sd, t['pixel'], t['rgb'], t['depth'], t['error'], t['points'])
