"""GPU: a whole training iteration captured as ONE CUDA graph (simple_rf_b200/train_graph.py) takes the same optimisation
steps as the eager loop: forward, reference-shaped losses, hand-written backward, fused Adam (capturable form: step count and
learning rate in device memory).  Randomness is switched off (no jitter, no sigma noise, white background) so that both
loops are deterministic functions of the batch; the learning rate is changed between iterations the way the trainer does
(src/Trainer10.py:303-308)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _setup(kind):
    from simple_rf_b200 import synthetic
    if kind == 'nerf':
        from simple_rf_b200.models.SimpleNeRF91 import SimpleNeRF as Model
        cfg = synthetic.nerf_configs(rng_mode='device')
        cfg['model']['raw_noise_std'] = 0.0
        mc = synthetic.scene_model_configs('llff', num_views=3)
        keys = ('rgb_coarse', 'rgb_fine', 'points_augmentation_rgb_coarse', 'views_augmentation_rgb_coarse')
    else:
        from simple_rf_b200.models.SimpleTensoRF91 import SimpleTensoRF as Model
        cfg = synthetic.tensorf_configs(num_voxels=48 ** 3, rng_mode='device')
        cfg['model']['augmentations'][0]['coarse_model']['num_voxels_initial'] = 24 ** 3
        cfg['model']['white_bkgd'] = True
        mc = synthetic.scene_model_configs('re10k', num_views=3)
        keys = ('rgb_coarse', 'points_augmentation_rgb_coarse')
    cfg['model']['perturb'] = False
    torch.manual_seed(0)
    model = Model(cfg, mc).to(DEV).train()
    if kind == 'tensorf':
        with torch.no_grad():
            for t in (model.coarse_model, model.augmented_models[0]['coarse_model']):
                for p in t.matrices_density:
                    p.mul_(6.0)
    ocfg = cfg['optimizers'][0]
    opt = torch.optim.Adam(model.get_trainable_parameters(ocfg), betas=(ocfg['beta1'], ocfg['beta2']))
    model.optimizers = {'optimizer_nerf': opt}
    h, w = mc['resolution']
    g = torch.Generator().manual_seed(4)
    n = 384
    pid = torch.stack([torch.randint(0, 3, (n,), generator=g), torch.randint(0, w, (n,), generator=g), torch.randint(0, h, (n,), generator=g)], 1).int().to(DEV)
    target = torch.rand(n, 3, generator=g).to(DEV)
    batch = {'pixel_id': pid, 'num_frames': 3, 'iter_num': 0, 'sub_batch_index': 0}

    def step():
        opt.zero_grad(set_to_none=True)
        out = model(batch)
        loss = sum(((out[k] - target) ** 2).mean() for k in keys) + 0.1 * (out['depth_coarse'] - out['points_augmentation_depth_coarse'].detach()).square().mean()
        loss.backward()
        opt.step()
        return loss.detach()
    return model, opt, step, batch


def _lr_schedule(opt, it):
    for group in opt.param_groups:
        group['lr'] = group['lr'] * (0.9 if it % 2 else 1.0)


@pytest.mark.parametrize('kind', ['nerf', 'tensorf'])
def test_graphed_iterations_equal_eager_iterations(kind):
    from simple_rf_b200.train_graph import GraphedStep
    iters, warm = 6, 2
    model_e, opt_e, step_e, _ = _setup(kind)
    losses_e = []
    for it in range(iters):
        losses_e.append(step_e().item())
        _lr_schedule(opt_e, it)
    model_g, opt_g, step_g, batch = _setup(kind)
    losses_g = []
    orig = step_g

    def step_logged():                       # the warm-up iterations inside GraphedStep are real optimisation steps
        out = orig()
        if not torch.cuda.is_current_stream_capturing():
            losses_g.append(out.item())
            _lr_schedule(opt_g, len(losses_g) - 1)
        return out
    graphed = GraphedStep(step_logged, {'optimizer_nerf': opt_g}, warmup=warm)
    assert len(losses_g) == warm
    for it in range(warm, iters):
        losses_g.append(graphed.replay().item())
        _lr_schedule(opt_g, it)
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= 2e-4 * abs(a), (losses_e, losses_g)
    pe, pg = dict(model_e.named_parameters()), dict(model_g.named_parameters())
    worst = 0.0
    for n_, p in pe.items():
        q = pg[n_]
        worst = max(worst, float((p.detach() - q.detach()).norm() / p.detach().norm().clamp_min(1e-8)))
    assert worst <= 2e-3, worst
    # host mirrors follow the replays: torch-format optimiser state reports the right step count
    steps = {float(st['step']) for st in opt_g.state.values() if 'step' in st}
    fused = opt_g._srf_fused
    assert all(s == iters for fg in fused.groups if fg is not None for s in fg.steps)
    assert int(fused.groups[0].step_dev.item()) == iters
    # derived caches (packed bf16 weights, channels-last planes) see the replayed updates: eval renders agree
    model_e.eval(); model_g.eval()
    with torch.no_grad():
        a = model_e({'pixel_id': batch['pixel_id'], 'num_frames': 3})
        b = model_g({'pixel_id': batch['pixel_id'], 'num_frames': 3})
    key = 'rgb_fine' if kind == 'nerf' else 'rgb_coarse'
    assert (a[key] - b[key]).abs().max().item() <= 5e-3
    print(kind, 'parameter drift', worst, 'losses', losses_g)
