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
def test_tensorf_dropin_resolves_and_matches_state_dict():
    from oracle import reference_harness as H
    from simple_rf_b200 import dropin
    get_model, _ = H.import_reference()
    dropin.install()
    configs, model_configs = H.load_configs(212, '00000')
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
