"""GPU parity: the drop-in SimpleNeRF class end to end against the committed outputs of the unmodified
reference model (tests/golden/nerf_{eval,train}.npz), same seeds, same parameters."""
import pytest
import torch

from oracle import fixtures as FX

pytestmark = pytest.mark.gpu
DEV = 'cuda'
# bf16-operand tensor-core MLP: the looser, stated tolerance of the parity ledger (measured ~3e-4 at init)
MAP_TOL = 3e-3
EXACT_TOL = 1e-6


def _model(golden_configs, seed):
    from simple_rf_b200.models.SimpleNeRF91 import SimpleNeRF
    configs, mc = golden_configs('nerf')
    model = SimpleNeRF(configs, mc)
    sets = FX.nerf_param_sets(configs, seed=seed)
    model.coarse_model.load_state_dict(sets['coarse_model'])
    model.fine_model.load_state_dict(sets['fine_model'])
    for aug, (_, _, params) in zip(model.augmented_models, sets['augmentations']):
        aug['coarse_model'].load_state_dict(params)
    return model.to(DEV), configs, mc


@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_dropin_forward_vs_reference_golden(golden, golden_configs, mode):
    g = golden(f'nerf_{mode}')
    model, configs, mc = _model(golden_configs, int(g['param_seed']))
    model.train(mode == 'train')
    torch.manual_seed(int(g['rng_seed']))
    with torch.no_grad():
        out = model({'pixel_id': g['pixel_id'].to(DEV), 'num_frames': 3, 'iter_num': 0, 'sub_batch_index': 0}, retraw=True)
    for k in ('rays_o', 'rays_d', 'rays_o_ndc', 'rays_d_ndc', 'view_dirs'):
        assert (out[k].cpu() - g[k]).abs().max().item() <= EXACT_TOL * max(1.0, g[k].abs().max().item()), k
    assert torch.equal(out['z_vals_coarse'].cpu(), g['z_vals_coarse'])          # same CPU RNG stream, same arithmetic
    worst = {}
    for k, ref in g.items():
        if k not in out or ref.dtype != torch.float32 or k.startswith('z_vals') or k.startswith('rays') or k == 'view_dirs':
            continue
        got = out[k].cpu()
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        err = (got - ref).abs().max().item() / max(1.0, ref.abs().max().item())
        worst[k] = err
        # depth_var of a nearly empty ray is ill-conditioned in fp32 world depths: compared relatively above
        assert err <= MAP_TOL, (k, err)
    assert (out['z_vals_fine'].cpu() - g['z_vals_fine']).abs().max().item() <= MAP_TOL
    print(mode, 'worst:', sorted(worst.items(), key=lambda kv: -kv[1])[:4])
    expect = {k for k in g if k in ('rgb_coarse', 'depth_fine', 'points_augmentation_rgb_coarse', 'views_augmentation_depth_coarse')}
    assert expect <= set(out.keys())


def test_dropin_eval_split_bf16_meets_the_fp32_contract(golden, golden_configs):
    """`configs['model']['mlp_precision'] = 'bf16x3'`: every map and per-sample tensor of the test-time render within 1e-3 of the
    unmodified fp32 reference (north_star's bound for fp32 paths), fine depths included."""
    g = golden('nerf_eval')
    model, configs, mc = _model(golden_configs, int(g['param_seed']))
    model.configs['model']['mlp_precision'] = 'bf16x3'
    model.eval()
    with torch.no_grad():
        out = model({'pixel_id': g['pixel_id'].to(DEV), 'num_frames': 3, 'iter_num': 0, 'sub_batch_index': 0}, retraw=True)
    worst = {}
    for k, ref in g.items():
        if k not in out or ref.dtype != torch.float32:
            continue
        worst[k] = (out[k].cpu() - ref).abs().max().item() / max(1.0, ref.abs().max().item())
        assert worst[k] <= 1e-3, (k, worst[k])
    print('bf16x3 worst:', sorted(worst.items(), key=lambda kv: -kv[1])[:5])
    assert {'rgb_fine', 'depth_fine', 'weights_fine', 'raw_sigma_coarse', 'z_vals_fine'} <= set(worst)


def test_dropin_retraw_false_drops_per_sample_outputs(golden, golden_configs):
    g = golden('nerf_eval')
    model, *_ = _model(golden_configs, int(g['param_seed']))
    model.eval()
    with torch.no_grad():
        out = model({'pixel_id': g['pixel_id'].to(DEV), 'num_frames': 3})
    assert 'weights_coarse' not in out and 'z_vals_fine' not in out and 'raw_sigma_fine' not in out
    for k in ('rgb_fine', 'depth_fine', 'depth_ndc_coarse', 'acc_fine', 'intrinsics', 'extrinsics', 'extrinsics_all'):
        assert k in out
    assert (out['rgb_fine'].cpu() - g['rgb_fine']).abs().max().item() <= MAP_TOL


def test_dropin_training_step_gradients(golden, golden_configs):
    """One optimiser-facing step: loss over rgb/depth of every model; parameter gradients against the fp32
    oracle pipeline differentiated by autograd.  Stated tolerance for the all-bf16-operand MLP forward + backward
    (tests/test_gpu_nerf_mlp_bwd.py pins the kernels against their own arithmetic model at 3e-2): relative L2 error
    <= 0.15 per tensor against fp32 autograd."""
    from oracle import pipeline as P
    g = golden('nerf_train')
    model, configs, mc = _model(golden_configs, int(g['param_seed']))
    model.train()
    pid = g['pixel_id']
    torch.manual_seed(int(g['rng_seed']))
    out = model({'pixel_id': pid.to(DEV), 'num_frames': 3, 'iter_num': 0, 'sub_batch_index': 0})
    keys = ['rgb_coarse', 'rgb_fine', 'depth_coarse', 'depth_fine', 'points_augmentation_rgb_coarse',
            'views_augmentation_depth_coarse', 'depth_ndc_coarse']
    loss = sum(out[k].square().mean() for k in keys)
    loss.backward()
    sets = FX.nerf_param_sets(configs, seed=int(g['param_seed']))
    leaves = []
    for p in [sets['coarse_model'], sets['fine_model']] + [a[2] for a in sets['augmentations']]:
        for k in p:
            p[k] = p[k].clone().requires_grad_()
            leaves.append(p[k])
    torch.manual_seed(int(g['rng_seed']))
    ref = P.nerf_render_chunk(sets, configs, mc, pid, training=True)
    ref_loss = sum(ref[k].square().mean() for k in keys)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 2e-3 * abs(ref_loss.item())
    mods = [('coarse_model', model.coarse_model, sets['coarse_model']), ('fine_model', model.fine_model, sets['fine_model'])]
    mods += [(a['name'], a['coarse_model'], s[2]) for a, s in zip(model.augmented_models, sets['augmentations'])]
    worst = 0.0
    for name, mod, ps in mods:
        for k, p in mod.named_parameters():
            gr = ps[k].grad
            assert p.grad is not None, (name, k)
            rel = ((p.grad.cpu() - gr).norm() / gr.norm().clamp_min(1e-12)).item()
            worst = max(worst, rel)
            assert rel <= 0.15, (name, k, rel)
    print('worst relative gradient error', worst)


@pytest.mark.parametrize('fused_adam', [True, False])
def test_training_curve_follows_reference(golden, golden_configs, fused_adam, monkeypatch):
    """Loss scalars per iteration (parity ledger): 8 iterations of forward + hand-written backward + Adam on the
    batches of tests/golden/nerf_train_curve.npz, against the curve the UNMODIFIED reference model produced with torch autograd
    and torch.optim.Adam on the CPU (oracle/generate_golden.py::golden_nerf_training_curve).  Same parameters, batches and RNG
    stream.  Stated tolerance for the bf16-operand tensor-core forward / backward: 0.1 % of the loss at every iteration
    (measured: 2e-5), parameter norms of the coarse MLP within 1e-3 relative after the last step (measured: 5e-4)."""
    monkeypatch.setenv('SIMPLE_RF_B200_FUSED_ADAM', '1' if fused_adam else '0')
    g = golden('nerf_train_curve')
    model, configs, mc = _model(golden_configs, int(g['param_seed']))
    model.train()
    opt_cfg = next(c for c in configs['optimizers'] if c['name'] == 'optimizer_main')
    opt = torch.optim.Adam(model.get_trainable_parameters(opt_cfg), lr=float(g['lr']), betas=(float(g['beta1']), float(g['beta2'])))
    model.optimizers = {'optimizer_nerf': opt}           # Trainer10.py:59-62
    assert bool(getattr(model, '_fused_adam', [])) == fused_adam
    rgb_keys = ('rgb_coarse', 'rgb_fine', 'points_augmentation_rgb_coarse', 'views_augmentation_rgb_coarse')
    torch.manual_seed(int(g['rng_seed']))
    worst = 0.0
    for it in range(g['loss'].shape[0]):
        pid, target = g['pixel_id'][it].to(DEV), g['target'][it].to(DEV)
        opt.zero_grad(set_to_none=True)
        out = model({'pixel_id': pid, 'num_frames': 3, 'iter_num': it, 'sub_batch_index': 0})
        loss = sum(((out[k] - target) ** 2).mean() for k in rgb_keys)
        loss = loss + 0.1 * (out['depth_coarse'] - out['points_augmentation_depth_coarse'].detach()).square().mean()
        loss.backward()
        opt.step()
        ref = g['loss'][it].item()
        rel = abs(loss.item() - ref) / ref
        worst = max(worst, rel)
        assert rel <= 1e-3, (it, loss.item(), ref)
    norms = torch.stack([p.detach().norm() for p in model.coarse_model.parameters()]).cpu()
    rel_n = ((norms - g['coarse_param_norms']).abs() / g['coarse_param_norms'].clamp_min(1e-6)).max().item()
    print(f'worst relative loss deviation {worst:.2e}, worst relative parameter-norm deviation {rel_n:.2e}')
    assert rel_n <= 1e-3


def test_camera_tables_follow_load_state_dict(golden, golden_configs):
    """The per-view camera tables the ray-generation kernel reads are a derived cache: a checkpoint loaded into a model that already
    rendered (`load_state_dict` copies new cameras in place) must move the rays (SimpleNeRF17.py:119-131 rebuilds per call)."""
    import copy
    g = golden('nerf_eval')
    model, configs, mc = _model(golden_configs, int(g['param_seed']))
    model.eval()
    pid = g['pixel_id'][:16].to(DEV)
    with torch.no_grad():
        before = model({'pixel_id': pid, 'num_frames': 3})
    state = copy.deepcopy(model.state_dict())
    key = 'extrinsics_learner.initial_extrinsics'
    assert key in state
    state[key][:, :3, 3] += 0.25                                  # translate every camera
    model.load_state_dict(state)
    with torch.no_grad():
        after = model({'pixel_id': pid, 'num_frames': 3})
    shift = (after['rays_o'] - before['rays_o']).abs().max().item()
    assert shift > 0.1, shift
    assert torch.equal(after['rays_d'], before['rays_d'])         # same rotations, same directions


@pytest.mark.parametrize('mode', ['eval', 'train'])
def test_dropin_variant_config_vs_reference_golden(golden, golden_configs, mode):
    """The switches no shipped run flips — world-space sampling (`data_loader.ndc = False`), `lindisp`, `white_bkgd`, another noise level —
    against the unmodified reference (tests/golden/nerf_variant_*.npz)."""
    from simple_rf_b200.models.SimpleNeRF91 import SimpleNeRF
    g = golden(f'nerf_variant_{mode}')
    configs, mc = golden_configs('nerf_variant')
    model = SimpleNeRF(configs, mc)
    sets = FX.nerf_param_sets(configs, seed=int(g['param_seed']))
    model.coarse_model.load_state_dict(sets['coarse_model'])
    model.fine_model.load_state_dict(sets['fine_model'])
    for aug, (_, _, params) in zip(model.augmented_models, sets['augmentations']):
        aug['coarse_model'].load_state_dict(params)
    model = model.to(DEV)
    model.train(mode == 'train')
    torch.manual_seed(int(g['rng_seed']))
    with torch.no_grad():
        out = model({'pixel_id': g['pixel_id'].to(DEV), 'num_frames': 3, 'iter_num': 0, 'sub_batch_index': 0}, retraw=True)
    assert not any('ndc' in k for k in out)
    for k in ('rays_o', 'rays_d', 'view_dirs'):
        assert (out[k].cpu() - g[k]).abs().max().item() <= EXACT_TOL * max(1.0, g[k].abs().max().item()), k
    assert torch.equal(out['z_vals_coarse'].cpu(), g['z_vals_coarse'])          # lindisp ladder + CPU jitter: same arithmetic
    worst = {}
    for k, ref in g.items():
        if k not in out or ref.dtype != torch.float32 or k.startswith('z_vals') or k.startswith('rays') or k == 'view_dirs':
            continue
        got = out[k].cpu()
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        worst[k] = (got - ref).abs().max().item() / max(1.0, ref.abs().max().item())
        # disparity-linear world depths put intervals of up to ~0.5 at the far end (0.02 in the NDC fixture): alpha = 1 - exp(-sigma * delta)
        # is that much more sensitive to the bf16 rounding of sigma.  Stated bf16 bound here: 1e-2 (the trained-field bound of DESIGN.md §5);
        # the fp32-contract program below must stay within 1e-3 on the same inputs, which separates rounding from a defect.
        assert worst[k] <= 1e-2, (k, worst[k])
    assert (out['z_vals_fine'].cpu() - g['z_vals_fine']).abs().max().item() <= 1e-2 * max(1.0, g['z_vals_fine'].abs().max().item())
    print('variant', mode, 'bf16 worst:', sorted(worst.items(), key=lambda kv: -kv[1])[:4])
    if mode == 'eval':
        model.configs['model']['mlp_precision'] = 'bf16x3'
        with torch.no_grad():
            out = model({'pixel_id': g['pixel_id'].to(DEV), 'num_frames': 3, 'iter_num': 0, 'sub_batch_index': 0}, retraw=True)
        worst = {}
        for k, ref in g.items():
            if k.endswith('_coarse') and ref.dtype == torch.float32 and k in out:          # coarse pass: same depths on both sides
                worst[k] = (out[k].cpu() - ref).abs().max().item() / max(1.0, ref.abs().max().item())
                assert worst[k] <= 1e-3, (k, worst[k])
        print('variant eval bf16x3 worst (coarse):', sorted(worst.items(), key=lambda kv: -kv[1])[:4])
