"""GPU: multi-rank training through the reference's unmodified `Trainer.train_one_iter` reproduces the single-process run.

Two processes share cuda:0 (NCCL refuses two ranks on one device, so the process group is `gloo`, which all-reduces CUDA
tensors through the host — the collective's *semantics* are what is under test; NCCL itself is exercised by `bench.py --gpus N`
and tools/train_ddp.py).  Covered: `DataPreprocessor91.select_batch_indices` rank slicing, the `srf_shard` replay of the
global CPU random stream inside the models, `FusedFlatAdam`'s single all-reduce of the flat gradient bucket with the
per-parameter 'has gradient' flags (parameters without a gradient on every rank must not be stepped: TensoRF leaves stale
planes in the optimiser between `shrink_tensor` and the next `reconfigure_optimizer`, SimpleTensoRF09.py:821-830), and the
test-time row-band sharding inside the drop-in's `render()` (SURVEY.md §8e iii)."""
import os
import socket

import numpy
import pytest
import torch
import torch.multiprocessing as mp

from simple_rf_b200.dropin import callers as C

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not C.available(), reason='upstream tree not installed (tools/install_reference.sh)')]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nerf_cfg():
    cfg = C.use_dropin(C.complete_configs(C.load_shipped_configs(1142), [0], seed=21))
    for loss in cfg['losses']:
        if 'iter_weights' in loss:
            loss['iter_weights'] = {'0': 0.1}
    cfg['data_loader']['num_rays'] = 1024
    cfg['data_loader']['sparse_depth']['num_rays'] = 1024
    return cfg, C.synthetic_raw_data('llff', 3, resolution=(126, 168), sparse_points=500, seed=9)


def _tensorf_cfg():
    cfg = C.use_dropin(C.complete_configs(C.load_shipped_configs(212), [0], seed=22))
    for loss in cfg['losses']:
        if 'iter_weights' in loss:
            loss['iter_weights'] = {'0': 0.01 if loss['name'].startswith('MassConcentration') else 0.1}
    for t in [cfg['model']['coarse_model']] + [a['coarse_model'] for a in cfg['model']['augmentations']]:
        t['tensor_upsampling_iters'] = [1, 6, 400, 550]
        t['alpha_mask_update_iters'] = [2]            # alpha mask + shrink at 2; optimiser re-grouped only at 6: stale planes for 4 steps
    cfg['model']['coarse_model']['num_voxels_initial'] = 64 ** 3
    cfg['model']['coarse_model']['num_voxels_final'] = 128 ** 3
    cfg['data_loader']['num_rays'] = 1024
    cfg['data_loader']['sparse_depth']['num_rays'] = 1024
    return cfg, C.synthetic_raw_data('re10k', 3, resolution=(96, 160), sparse_points=500, seed=10, tensorf=True)


def _train(kind, iters, world, rank):
    cfg, raw = _nerf_cfg() if kind == 'nerf' else _tensorf_cfg()
    trainer, model, mc = C.make_trainer(cfg, raw, seed=cfg['seed'])
    curve, stale_steps = [], []
    opt = trainer.optimizers['optimizer_nerf']
    for it in range(iters):
        losses = trainer.train_one_iter(it)
        C.step_learning_rates(trainer, it)
        curve.append({k: float(v) for k, v in losses.items()})
        live = {id(p) for p in model.module.parameters()}
        stale_steps.append(sorted(float(st['step']) for p, st in opt.state.items() if id(p) not in live and 'step' in st))
    torch.cuda.synchronize()
    params = {n: p.detach().float().cpu() for n, p in model.module.named_parameters()}
    # test-time frame through the model's own band sharding
    model.eval()
    h, w = mc['resolution']
    ys, xs = numpy.meshgrid(numpy.arange(0, h, 3, dtype=numpy.int32), numpy.arange(0, w, 3, dtype=numpy.int32), indexing='ij')
    pid = torch.from_numpy(numpy.stack([numpy.zeros(xs.size, dtype=numpy.int32), xs.reshape(-1), ys.reshape(-1)], 1)).cuda()
    with torch.no_grad():
        out = model.module({'pixel_id': pid, 'num_frames': 3})
    key = 'rgb_fine' if kind == 'nerf' else 'rgb_coarse'
    frame = {k: out[k].float().cpu() for k in (key, key.replace('rgb', 'depth'), 'rays_d')}
    return curve, params, stale_steps, frame


def _worker(rank, world, port, kind, iters, results):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK='0')
    torch.cuda.set_device(0)
    from simple_rf_b200 import parallel
    parallel.init_from_env(backend='gloo')
    curve, params, stale, frame = _train(kind, iters, world, rank)
    results.put((rank, curve, {k: v.numpy() for k, v in params.items()}, stale, {k: v.numpy() for k, v in frame.items()}))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def _run_world2(kind, iters):
    ctx = mp.get_context('spawn')
    results = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, iters, results)) for r in range(2)]
    for p in procs:
        p.start()
    got = [results.get() for _ in procs]
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    return sorted(got, key=lambda g: g[0])


def _check(kind, iters, loss_tol, drift_tol):
    single_curve, single_params, single_stale, single_frame = _train(kind, iters, 1, 0)
    ranks = _run_world2(kind, iters)
    # ranks end bit-identical: same averaged gradient, same fused step
    for n in single_params:
        assert numpy.array_equal(ranks[0][2][n], ranks[1][2][n]), n
    # mean of the two ranks' loss means == the single-process mean (equal shares of both ray kinds)
    worst = 0.0
    for it in range(iters):
        for k, v in single_curve[it].items():
            if k.startswith('TotalVariation') or k.startswith('lr_'):
                both = ranks[0][1][it][k]                          # computed identically on every rank
            else:
                both = 0.5 * (ranks[0][1][it][k] + ranks[1][1][it][k])
            rel = abs(both - v) / max(abs(v), 1e-8)
            worst = max(worst, rel)
            assert rel <= loss_tol, (it, k, v, both)
    drift = max(float(numpy.linalg.norm(ranks[0][2][n] - single_params[n].numpy()) / max(float(single_params[n].norm()), 1e-8))
                for n in single_params)
    assert drift <= drift_tol, drift
    # stale optimiser entries (no gradient on any rank) keep their step counters exactly as in the single-process run
    assert ranks[0][3] == single_stale and ranks[1][3] == single_stale
    # the band-sharded test-time frame equals the single-process frame on both ranks
    for r in ranks:
        for k, v in single_frame.items():
            assert r[4][k].shape == tuple(v.shape)
            assert float(numpy.abs(r[4][k] - v.numpy()).max()) <= 1e-3 + 10 * drift_tol, k
    print(kind, 'worst loss deviation', worst, 'parameter drift', drift, 'stale steps', single_stale[-1][:4])


def test_nerf_two_ranks_reproduce_single_process_training():
    _check('nerf', iters=3, loss_tol=2e-4, drift_tol=2e-3)


def test_tensorf_two_ranks_reproduce_single_process_training_across_shrink():
    _check('tensorf', iters=6, loss_tol=2e-3, drift_tol=5e-3)
