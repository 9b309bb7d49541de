"""CPU suite: the drop-in classes resolve through the reference's unmodified factory and expose the same
parameter names / shapes / optimiser groups as the reference models (checkpoint interchange)."""
import copy

import pytest
import torch


def _ref_and_dropin(name, train_num, scene, new_name):
    from oracle import reference_harness as H
    from simple_rf_b200 import dropin
    get_model, _ = H.import_reference()
    dropin.install()
    configs, model_configs = H.load_configs(train_num, scene)
    configs['model']['name'] = name
    torch.manual_seed(0)
    ref = get_model(copy.deepcopy(configs), model_configs=model_configs)
    configs['model']['name'] = new_name
    torch.manual_seed(0)
    mine = get_model(copy.deepcopy(configs), model_configs=model_configs)
    return ref, mine, configs


@pytest.mark.needs_reference
def test_nerf_dropin_resolves_and_matches_state_dict():
    ref, mine, configs = _ref_and_dropin('SimpleNeRF17', 1142, 'fern', 'SimpleNeRF91')
    assert type(mine).__module__.startswith('simple_rf_b200.models') and type(mine).__name__ == 'SimpleNeRF'
    sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
    assert list(sd_ref.keys()) == list(sd_mine.keys())
    for k in sd_ref:
        assert sd_ref[k].shape == sd_mine[k].shape, k
        assert torch.equal(sd_ref[k], sd_mine[k]), k            # same construction order => same seeded init
    mine.load_state_dict(sd_ref)
    opt_cfg = configs['optimizers'][0]
    g_ref, g_mine = ref.get_trainable_parameters(opt_cfg), mine.get_trainable_parameters(opt_cfg)
    assert [g['name'] for g in g_ref] == [g['name'] for g in g_mine]
    assert [len(list(g['params'])) for g in g_ref] == [len(list(g['params'])) for g in g_mine]
    assert [a['name'] for a in mine.augmented_models] == [a['name'] for a in ref.augmented_models]


def test_nerf_dropin_refuses_cpu(golden_configs):
    from simple_rf_b200._lib import SimpleRFNativeError
    from simple_rf_b200.models.SimpleNeRF91 import SimpleNeRF
    configs, mc = golden_configs('nerf')
    model = SimpleNeRF(configs, mc).eval()
    pid = torch.zeros(4, 3, dtype=torch.int32)
    with pytest.raises(SimpleRFNativeError):
        model({'pixel_id': pid, 'num_frames': 3})


@pytest.mark.needs_reference
@pytest.mark.parametrize('decomposition', ['VectorMatrix', 'CandecompParafac'])
def test_tensorf_dropin_resolves_and_matches_state_dict(decomposition):
    from oracle import reference_harness as H
    from simple_rf_b200 import dropin
    get_model, _ = H.import_reference()
    dropin.install()
    configs, model_configs = H.load_configs(212, '00000')
    if decomposition == 'CandecompParafac':                    # SimpleTensoRF09.py:537-539; one component count per tensor (:992, :986)
        for cfg in (configs['model']['coarse_model'], configs['model']['augmentations'][0]['coarse_model']):
            cfg.update(decomposition_type=decomposition, num_components_density=[24], num_components_color=[48])
    configs['model']['coarse_model']['num_voxels_initial'] = 30 ** 3
    configs['model']['augmentations'][0]['coarse_model']['num_voxels_initial'] = 16 ** 3
    models = []
    for name in ('SimpleTensoRF09', 'SimpleTensoRF91'):
        configs['model']['name'] = name
        torch.manual_seed(0)
        models.append(get_model(copy.deepcopy(configs), model_configs=model_configs))
    ref, mine = models
    assert type(mine).__module__.startswith('simple_rf_b200.models')
    sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
    assert list(sd_ref.keys()) == list(sd_mine.keys())
    for k in sd_ref:
        assert sd_ref[k].shape == sd_mine[k].shape, k
        assert torch.equal(sd_ref[k], sd_mine[k]), k
    opt_cfg = configs['optimizers'][0]
    g_ref, g_mine = ref.get_trainable_parameters(opt_cfg), mine.get_trainable_parameters(opt_cfg)
    assert [g['name'] for g in g_ref] == [g['name'] for g in g_mine]
    assert [[tuple(p.shape) for p in g['params']] for g in g_ref] == [[tuple(p.shape) for p in g['params']] for g in g_mine]
    assert int(mine.coarse_model.num_samples) == int(ref.coarse_model.num_samples)
    assert type(mine.coarse_model).__name__ == type(ref.coarse_model).__name__          # TotalVariationLoss04.py:87 dispatches on it


def test_callers_harness_drives_unmodified_trainer_and_tester_on_cpu():
    """simple_rf_b200/dropin/callers.py (used by tests/test_gpu_reference_callers.py and bench.py --impl reference): the
    UNMODIFIED Trainer.train_one_iter (src/Trainer10.py:65) and NerfTester.predict_frame (src/Tester07.py:153) run on the
    analytic synthetic scene with the reference's own classes, on the CPU, at a seconds-scale size."""
    from simple_rf_b200.dropin import callers as C
    if not C.available():
        pytest.skip('upstream tree not present')
    import numpy
    cfg = C.complete_configs(C.load_shipped_configs(1142), [0], seed=3)
    cfg['data_loader']['num_rays'] = 64
    cfg['data_loader']['sparse_depth']['num_rays'] = 64
    raw = C.synthetic_raw_data('llff', 3, resolution=(30, 40), sparse_points=100, seed=2)
    # the scene is consistent: a sparse depth re-projects onto the same colour in the neighbouring view
    assert raw['nerf_data']['images'].shape == (3, 30, 40, 3) and raw['nerf_data']['images'].dtype == numpy.uint8
    trainer, model, mc = C.make_trainer(cfg, raw, seed=3)
    assert mc['resolution'] == [30, 40] and abs(mc['near'] - 1.0) < 1e-6 and mc['near_ndc'] == 0.0
    losses = trainer.train_one_iter(0)
    assert set(losses) == {'MSE14', 'SparseDepthMSE14', 'AugmentationsDepthLoss11', 'CoarseFineConsistencyLoss34', 'TotalLoss'}
    assert all(numpy.isfinite(v) for v in losses.values())
    tester = C.make_tester(cfg, mc, [0])
    tester.model.load_state_dict(model.state_dict())
    tester.model.eval()
    frame = tester.predict_frame(C.test_pose(raw))
    assert frame['image'].shape == (30, 40, 3) and frame['image'].dtype == numpy.uint8
    assert set(frame) == {'image', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc'}
    names = C.use_dropin(cfg)
    assert names['model']['name'] == 'SimpleNeRF91' and names['data_loader']['data_preprocessor_name'] == 'DataPreprocessor91'
    assert [l['name'] for l in names['losses']] == ['MSE14', 'SparseDepthMSE14', 'AugmentationsDepthLoss91', 'CoarseFineConsistencyLoss91']
