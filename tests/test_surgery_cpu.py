"""CPU: grid-surgery oracle (oracle/surgery.py) against the committed outputs of the unmodified reference, and the product's
host-side surgery logic (simple_rf_b200/grid_surgery.py: plan, crop window, optimiser re-grouping) against the oracle."""
import copy

import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import surgery as SG
from oracle import tensorf as TF
from simple_rf_b200 import grid_surgery as GS

_bits, run_oracle_schedule = SG.pack_volume, SG.replay_golden_schedule


def test_oracle_reproduces_reference_surgery(golden, golden_configs):
    g = golden('tensorf_surgery')
    configs, _ = golden_configs('tensorf_surgery')
    o = run_oracle_schedule(configs)
    assert torch.equal(_bits(o['vol1']), g['volume1_bits']) and list(o['vol1'].shape) == g['volume1_shape'].tolist()
    assert torch.equal(o['box1'], g['box1'])
    assert torch.equal(o['lo'], g['window_lo']) and torch.equal(o['hi'], g['window_hi'])
    assert torch.equal(o['geo1']['resolution'], g['shrink_resolution']) and torch.equal(o['geo1']['bbox'], g['shrink_bbox'])
    assert o['geo1']['num_samples'] == int(g['shrink_num_samples'])
    assert torch.equal(o['geo2']['resolution'], g['upsample_resolution']) and o['geo2']['num_samples'] == int(g['upsample_num_samples'])
    assert torch.equal(o['params2']['matrices_density.0'], g['upsampled_matrices_density_0'])
    assert torch.equal(o['params2']['vectors_color.2'], g['upsampled_vectors_color_2'])
    assert torch.equal(_bits(o['vol2']), g['volume2_bits']) and torch.equal(o['box2'], g['box2'])
    assert 0.01 < float(g['occupied1']) < 0.5 and float(g['occupied2']) > 0       # the fixture is neither empty nor full


def test_oracle_reproduces_reference_surgery_cp(golden, golden_configs):
    """The same schedule on a CANDECOMP/PARAFAC tensor (lines only: SimpleTensoRF09.py:1094-1124)."""
    g = golden('tensorf_cp_surgery')
    configs, _ = golden_configs('tensorf_cp_surgery')
    o = run_oracle_schedule(configs)
    assert TF.is_cp(o['params0'])
    assert torch.equal(_bits(o['vol1']), g['volume1_bits']) and list(o['vol1'].shape) == g['volume1_shape'].tolist()
    assert torch.equal(o['box1'], g['box1'])
    assert torch.equal(o['lo'], g['window_lo']) and torch.equal(o['hi'], g['window_hi'])
    assert torch.equal(o['geo1']['resolution'], g['shrink_resolution']) and torch.equal(o['geo1']['bbox'], g['shrink_bbox'])
    assert torch.equal(o['geo2']['resolution'], g['upsample_resolution']) and o['geo2']['num_samples'] == int(g['upsample_num_samples'])
    assert torch.equal(o['params2']['vectors_density.0'], g['upsampled_vectors_density_0'])
    assert torch.equal(o['params2']['vectors_color.2'], g['upsampled_vectors_color_2'])
    assert torch.equal(_bits(o['vol2']), g['volume2_bits']) and torch.equal(o['box2'], g['box2'])
    assert 0.01 < float(g['occupied1']) < 0.5 and float(g['occupied2']) > 0


def test_plan_matches_the_reference_schedule():
    tc = {'tensor_upsampling_iters': [2000, 3000, 4000, 5500], 'alpha_mask_update_iters': [2500, 4000],
          'num_voxels_initial': 2097156, 'num_voxels_final': 262144000}
    plan = GS.build_plan(tc)
    assert sorted(plan) == [2000, 2500, 3000, 4000, 5500]
    for it in tc['tensor_upsampling_iters']:
        want = SG.new_num_voxels(it, tc['tensor_upsampling_iters'], tc['num_voxels_initial'], tc['num_voxels_final'])
        assert plan[it][-1] == ('resample', want)
    assert plan[2500] == (('occupancy', True),)                       # crop with the first rebuild only
    assert plan[4000][0] == ('occupancy', False) and plan[4000][1][0] == 'resample'       # rebuild BEFORE the upsampling (:822-829)
    assert GS.voxel_ladder(tc)[5500] == tc['num_voxels_final']
    assert GS.build_plan({**tc, 'alpha_mask_update_iters': []})[4000] == (('resample', plan[4000][1][1]),)


@pytest.mark.parametrize('alpha_res_differs', [False, True])
def test_crop_window_matches_oracle(golden_configs, alpha_res_differs):
    configs, _ = golden_configs('tensorf_surgery')
    cfg = configs['model']['coarse_model']
    g = torch.Generator().manual_seed(5)
    for trial in range(20):
        res = torch.randint(20, 90, (3,), generator=g)
        half = 0.5 + 2 * torch.rand(3, generator=g)
        bbox = torch.stack([-half, half * (0.7 + 0.6 * torch.rand(3, generator=g))])
        geo = SG.tensor_geometry(res, bbox, cfg['num_voxels_per_sample'], cfg['num_samples_max'])
        a, b = torch.rand(3, generator=g) * 0.4, 0.6 + torch.rand(3, generator=g) * 0.4
        new_box = torch.stack([bbox[0] + a * geo['size'], bbox[0] + b * geo['size']])
        alpha_res = res + 3 if alpha_res_differs else res.clone()
        lo, hi, box = SG.shrink_window(geo, new_box, alpha_res)
        lo2, hi2, box2 = GS.crop_window(geo['bbox'], geo['voxel_length'], geo['resolution'], new_box, alpha_res)
        assert torch.equal(lo, lo2) and torch.equal(hi, hi2) and torch.equal(box, box2), trial


def _toy_optimizer(stepped):
    """Three named groups; only the parameters listed in `stepped` receive a gradient (and therefore a state entry)."""
    torch.manual_seed(0)
    groups = [{'name': 'a_tensor_params', 'params': [torch.nn.Parameter(torch.randn(3)) for _ in range(3)], 'lr': 0.02},
              {'name': 'a_network_params', 'params': [torch.nn.Parameter(torch.randn(2)) for _ in range(2)], 'lr': 0.001},
              {'name': 'b_tensor_params', 'params': [torch.nn.Parameter(torch.randn(4)) for _ in range(2)], 'lr': 0.02}]
    opt = torch.optim.Adam(groups)
    flat = [p for g in groups for p in g['params']]
    for i in stepped:
        flat[i].grad = torch.ones_like(flat[i])
    opt.step()
    for g in opt.param_groups:
        g['lr'] *= 0.5                                     # the trainer's decay (Trainer10.py:303-308)
    return opt, flat


@pytest.mark.parametrize('stepped', [range(7), [0, 1, 2, 3, 4], [1, 3, 5, 6], [], [0, 6]])
@pytest.mark.parametrize('which', [('a_tensor_params', 'a_network_params'), ('b_tensor_params',), ('a_network_params',)])
def test_regroup_optimizer_matches_the_reference_algorithm(stepped, which):
    outcomes = []
    for fn in (SG.reconfigure_optimizer, GS.regroup_optimizer):
        opt, flat = _toy_optimizer(list(stepped))
        fresh = [{'name': n, 'params': [torch.nn.Parameter(torch.zeros(5)) for _ in range(2)], 'lr': 0.3} for n in which]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            fn(opt, fresh)
        kept = [i for i, p in enumerate(flat) if p in opt.state]
        outcomes.append(([(g['name'], g['lr'], len(g['params'])) for g in opt.param_groups], kept))
    assert outcomes[0] == outcomes[1]


@pytest.mark.needs_reference
def test_regroup_restatement_is_the_reference_method():
    """oracle.surgery.reconfigure_optimizer == LowRankTensor.reconfigure_optimizer of the unmodified reference."""
    from oracle import generate_golden as G
    from oracle import reference_harness as H
    configs, model_configs = G.surgery_configs()
    opt_cfg = next(c for c in configs['optimizers'] if c['name'] == 'optimizer_main')
    results = []
    for use_reference in (True, False):
        torch.manual_seed(3)
        model = H.build_model(copy.deepcopy(configs), model_configs)
        opt = torch.optim.Adam(model.get_trainable_parameters(opt_cfg), betas=(opt_cfg['beta1'], opt_cfg['beta2']))
        model.optimizers = {'optimizer_nerf': opt}
        flat = [p for g in opt.param_groups for p in g['params']]
        for p in flat[::2]:
            p.grad = torch.zeros_like(p)
        opt.step()
        t = model.coarse_model
        if use_reference:
            t.reconfigure_optimizer()
        else:
            SG.reconfigure_optimizer(opt, t.get_trainable_parameters(opt_cfg))
        results.append(([(g['name'], g['lr'], len(g['params'])) for g in opt.param_groups], [i for i, p in enumerate(flat) if p in opt.state]))
    assert results[0] == results[1]
