"""GPU parity of the learnable-camera path (SimpleNeRF17.py:817-842 `ExtrinsicsLearner` with learn_camera_rotation / learn_camera_translation;
Tester07.py:62-111 refines test poses with it; no shipped config turns it on): the gradient of a fused-MLP evaluation w.r.t. its rays
(`srf_nerf_mlp_input_grad`) against fp32 autograd through the oracle MLP, and the pose-correction gradients r.grad / t.grad of the whole
drop-in model against the UNMODIFIED reference model driven by the unmodified trainer on the same box."""
import copy
import ctypes

import pytest
import torch

from oracle import nerf_mlp as M

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _variants(golden_configs):
    configs, mc = golden_configs('nerf')
    m = configs['model']
    return {'main': m['coarse_model'], 'points_augmentation': m['augmentations'][0]['coarse_model'],
            'views_augmentation': m['augmentations'][1]['coarse_model']}


def test_source_struct_matches_c_abi():
    from simple_rf_b200 import _lib, nerf_program
    assert _lib.load().srf_input_grad_source_bytes() == ctypes.sizeof(nerf_program.InputGradSource)


@pytest.mark.parametrize('variant', ['main', 'points_augmentation', 'views_augmentation'])
def test_mlp_input_gradients_vs_fp32_autograd(golden_configs, variant):
    from simple_rf_b200 import nerf_program as NP
    cfg = _variants(golden_configs)[variant]
    g = torch.Generator().manual_seed(91)
    params = M.init_mlp_params(cfg, g)
    params['pts_output_linear.bias'][0] += 1.0
    R, S = 41, 64                                   # 2624 rows: 20 full tiles + a ragged one
    o = (torch.rand(R, 3, generator=g) - .5).requires_grad_()
    d = (torch.rand(R, 3, generator=g) - .5).requires_grad_()
    vd = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).requires_grad_()
    z = torch.rand(R, S, generator=g)
    pts = (o[:, None] + d[:, None] * z[..., None]).reshape(-1, 3)
    vflat = vd[:, None].expand(R, S, 3).reshape(-1, 3)
    ref = M.mlp_forward(params, cfg, pts, vflat if cfg['use_view_dirs'] else None, None)
    g_sigma = torch.randn(R, S, 1, generator=g) * 0.1
    g_rgb = torch.randn(R, S, 3, generator=g)
    ((ref['sigma'] * g_sigma.reshape(-1, 1)).sum() + (ref['rgb'] * g_rgb.reshape(-1, 3)).sum()).backward()

    packed = NP.PackedMLP(cfg).refresh({k: v.to(DEV) for k, v in params.items()})
    od, dd, zd, vdd = o.detach().to(DEV), d.detach().to(DEV), z.to(DEV), vd.detach().to(DEV)
    sigma, rgb, acts = packed.forward(od, dd, zd, vdd if cfg['use_view_dirs'] else None, save=True)
    _, dz = NP.mlp_backward(packed, packed.flat, acts, sigma, rgb, g_sigma.to(DEV), g_rgb.to(DEV))
    g_o, g_d, g_v = NP.mlp_input_backward(packed, packed.flat, dz, od, dd, zd, vdd if cfg['use_view_dirs'] else None)
    torch.cuda.synchronize()
    report = {}
    for name, got, want in (('rays_o', g_o, o.grad), ('rays_d', g_d, d.grad), ('view_dirs', g_v, vd.grad if cfg['use_view_dirs'] else None)):
        if want is None:
            assert got is None
            continue
        rel = float((got.cpu() - want).norm() / want.norm())
        report[name] = rel
        # bf16 operands in the dgrad chain + the few ReLU masks a bf16 forward flips: same bound as the weight gradients
        assert rel <= 0.15, (name, rel)
    print(f'{variant}: input-gradient rel-L2 {report}')


def test_learnable_cameras_pose_gradients_vs_reference_trainer():
    """One `Trainer.train_one_iter` (src/Trainer10.py:65-115, unmodified) with learn_camera_rotation / learn_camera_translation on and a
    non-zero pose correction: losses and EVERY gradient incl. `extrinsics_learner.r` / `.t` against the reference's eager autograd."""
    from simple_rf_b200.dropin import callers as C
    if not C.available():
        pytest.skip('upstream tree not installed (tools/install_reference.sh)')
    from test_gpu_reference_callers import _compare_curves, _compare_grads, _nerf_configs, _report, _run_trainer
    raw = C.synthetic_raw_data('llff', 3, resolution=(126, 168), sparse_points=400, seed=6)
    cfg_ref = _nerf_configs()
    cfg_ref['model']['learn_camera_rotation'] = True
    cfg_ref['model']['learn_camera_translation'] = True
    cfg_mine = C.use_dropin(cfg_ref)

    def perturb(module):
        g = torch.Generator().manual_seed(5)
        lr = module.extrinsics_learner
        lr.r.data.copy_((torch.randn(lr.r.shape, generator=g) * 0.01).to(lr.r.device))
        lr.t.data.copy_((torch.randn(lr.t.shape, generator=g) * 0.02).to(lr.t.device))

    ref_curve, ref_grads, _, _, _ = _run_trainer(cfg_ref, raw, 1, prepare=perturb)
    my_curve, my_grads, my_model, _, _ = _run_trainer(cfg_mine, raw, 1, prepare=perturb)
    assert type(my_model.module).__module__.startswith('simple_rf_b200.models')
    for name in ('extrinsics_learner.r', 'extrinsics_learner.t'):
        assert name in ref_grads and float(ref_grads[name].norm()) > 0, name
    worst = _compare_curves(ref_curve, my_curve, tol=2e-3)
    rels = _compare_grads(ref_grads, my_grads, tol=0.15)
    _report('nerf_learnable_cameras', {'loss_deviation': worst, 'gradient_relative_l2': rels,
                                       'r_grad_reference': ref_grads['extrinsics_learner.r'].tolist(),
                                       'r_grad_dropin': my_grads['extrinsics_learner.r'].tolist(),
                                       't_grad_reference': ref_grads['extrinsics_learner.t'].tolist(),
                                       't_grad_dropin': my_grads['extrinsics_learner.t'].tolist()})
    print('pose-correction gradients rel-L2:', {k: v for k, v in rels.items() if 'extrinsics' in k}, 'worst', max(rels.values()))


def test_tensorf_learnable_cameras_pose_gradients_vs_reference_trainer():
    """Simple-TensoRF (NDC): the pose correction receives its gradient through the view directions of the colour MLP, |d| under delta and the
    NDC -> world depths (grid coordinates are detached upstream, SimpleTensoRF09.py:1054, :1075).  One unmodified `train_one_iter`, every
    gradient incl. r / t against the reference's eager autograd."""
    from simple_rf_b200.dropin import callers as C
    if not C.available():
        pytest.skip('upstream tree not installed (tools/install_reference.sh)')
    from test_gpu_reference_callers import _compare_curves, _compare_grads, _report, _run_trainer, _tensorf_configs
    raw = C.synthetic_raw_data('re10k', 3, resolution=(144, 256), sparse_points=600, seed=9, tensorf=True)
    cfg_ref = _tensorf_configs()
    cfg_ref['model']['learn_camera_rotation'] = True
    cfg_ref['model']['learn_camera_translation'] = True
    cfg_mine = C.use_dropin(cfg_ref)

    def perturb(module):
        g = torch.Generator().manual_seed(15)
        lr = module.extrinsics_learner
        lr.r.data.copy_((torch.randn(lr.r.shape, generator=g) * 0.01).to(lr.r.device))
        lr.t.data.copy_((torch.randn(lr.t.shape, generator=g) * 0.02).to(lr.t.device))

    ref_curve, ref_grads, _, _, _ = _run_trainer(cfg_ref, raw, 1, prepare=perturb)
    my_curve, my_grads, my_model, _, _ = _run_trainer(cfg_mine, raw, 1, prepare=perturb)
    assert type(my_model.module).__module__.startswith('simple_rf_b200.models')
    for name in ('extrinsics_learner.r', 'extrinsics_learner.t'):
        assert name in ref_grads and float(ref_grads[name].norm()) > 0, name
    worst = _compare_curves(ref_curve, my_curve, tol=5e-3)
    rels = _compare_grads(ref_grads, my_grads, tol=0.1)
    _report('tensorf_learnable_cameras', {'loss_deviation': worst, 'gradient_relative_l2': rels,
                                          'r_grad_reference': ref_grads['extrinsics_learner.r'].tolist(),
                                          'r_grad_dropin': my_grads['extrinsics_learner.r'].tolist(),
                                          't_grad_reference': ref_grads['extrinsics_learner.t'].tolist(),
                                          't_grad_dropin': my_grads['extrinsics_learner.t'].tolist()})
    print('pose-correction gradients rel-L2:', {k: v for k, v in rels.items() if 'extrinsics' in k}, 'worst', max(rels.values()))


def _refine_pose(cfg, mc, state, raw, seed, iterations):
    """The reference's UNMODIFIED `NerfTester.optimize_test_camera_params` (src/Tester07.py:62-125) on view 1 of the scene, starting from a
    perturbed pose: -> (pose before, pose after)."""
    import numpy
    from simple_rf_b200.dropin import callers as C
    C.prepare()
    import Trainer10
    tester = C.make_tester(cfg, mc, cfg['device'])
    tester.model.load_state_dict(state)
    tester.test_configs['optimize_camera_params'] = {'num_iterations': iterations}
    nd = raw['nerf_data']
    poses, ks = numpy.asarray(nd['extrinsics'], dtype=numpy.float64), numpy.asarray(nd['intrinsics'], dtype=numpy.float64)
    noisy = poses.copy()
    c, s = numpy.cos(0.02), numpy.sin(0.02)
    noisy[1] = noisy[1] @ numpy.array([[c, -s, 0, 0.03], [s, c, 0, -0.02], [0, 0, 1, 0.01], [0, 0, 0, 1]])
    Trainer10.init_seeds(seed)
    _, after = tester.optimize_test_camera_params(numpy.asarray(nd['images'])[1], ks[1], poses[1], poses, noisy)
    return after


def _pose_refinement_configs(device):
    from test_gpu_reference_callers import _nerf_configs
    cfg = _nerf_configs()
    cfg['device'] = device
    cfg['model']['learn_camera_rotation'] = True
    cfg['model']['learn_camera_translation'] = True
    main = next(o for o in cfg['optimizers'] if o['name'] == 'optimizer_main')
    cfg['optimizers'].append(dict(main, name='optimizer_extrinsics', lr_initial=1e-3))
    return cfg


def test_test_time_pose_refinement_through_unmodified_tester():
    """`NerfTester.optimize_test_camera_params` (src/Tester07.py:62-125: test-optimization preprocessor, `rebuild_camera_params_learners`,
    whole-image batches in `mode='test_camera_params_optimization'`, an Adam optimiser over r / t, `mode='camera_params_only'` read-back)
    drives the drop-in and the reference model from the same briefly trained weights.  Adam's first steps are sign steps of size lr, so
    the two refined poses are compared through the direction they moved in and the worst-case bound of a flipped sign."""
    import numpy
    from simple_rf_b200.dropin import callers as C
    if not C.available():
        pytest.skip('upstream tree not installed (tools/install_reference.sh)')
    from test_gpu_reference_callers import _report, _run_trainer
    raw = C.synthetic_raw_data('llff', 3, resolution=(126, 168), sparse_points=400, seed=6)
    cfg_ref = _pose_refinement_configs([0])
    cfg_mine = C.use_dropin(cfg_ref)
    _, _, model, mc, _ = _run_trainer(cfg_mine, raw, 60)                 # a field with some structure, trained by the drop-in
    state = copy.deepcopy(model.state_dict())
    iterations, lr = 4, 1e-3
    after_ref = _refine_pose(copy.deepcopy(cfg_ref), mc, state, raw, 21, iterations)
    after_mine = _refine_pose(copy.deepcopy(cfg_mine), mc, state, raw, 21, iterations)
    before = _refine_pose(copy.deepcopy(cfg_mine), mc, state, raw, 21, 0)
    move_ref, move_mine = (after_ref - before)[:3].reshape(-1), (after_mine - before)[:3].reshape(-1)
    assert float(numpy.abs(move_ref).max()) > 0.5 * lr and float(numpy.abs(move_mine).max()) > 0.5 * lr, 'the pose did not move'
    cosine = float(move_ref @ move_mine / (numpy.linalg.norm(move_ref) * numpy.linalg.norm(move_mine)))
    worst = float(numpy.abs(after_ref - after_mine).max())
    _report('nerf_pose_refinement', {'before': before.tolist(), 'after_reference': after_ref.tolist(), 'after_dropin': after_mine.tolist(),
                                     'cosine_of_moves': cosine, 'max_abs_difference': worst})
    print('pose refinement: cosine of the two moves', cosine, 'max |difference|', worst)
    assert cosine >= 0.7 and worst <= 2.5 * lr * iterations, (cosine, worst)
