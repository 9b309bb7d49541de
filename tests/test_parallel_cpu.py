"""CPU suite: the multi-process host logic (ray sharding, flat-bucket gradient all-reduce) with world_size 2
over gloo on 127.0.0.1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simple_rf_b200 import parallel


def test_shard_bounds_partition():
    for n in (0, 1, 7, 4096, 762048):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batch_keeps_proportions():
    mask = torch.zeros(4096, dtype=torch.bool)
    mask[:2048] = True                       # 2048 image rays + 2048 sparse-depth rays (train1142 batch)
    seen = []
    for r in range(8):
        rows = parallel.shard_batch(mask, r, 8)
        assert rows.numel() == 512 and int(mask[rows].sum()) == 256
        seen.append(rows)
    assert torch.equal(torch.sort(torch.cat(seen))[0], torch.arange(4096))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    parallel.init_from_env(backend='gloo')
    torch.manual_seed(0)                                  # replicated parameters
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 2))
    frozen = torch.nn.Parameter(torch.ones(3), requires_grad=False)
    opt = torch.optim.Adam([{'params': list(net.parameters()) + [frozen], 'name': 'g'}], lr=1e-2)
    hooks = parallel.attach_gradient_allreduce({'optimizer_main': opt})
    assert len(hooks) == 1
    g = torch.Generator().manual_seed(100)
    x_all = torch.randn(8, 5, generator=g)
    s, e = parallel.shard_bounds(8, rank, world)
    loss = net(x_all[s:e]).square().mean()               # rank-local mean over an equal share
    loss.backward()
    opt.step()                                            # pre-hook all-reduces the flat bucket
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        # reference: single process on the whole batch
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 2))
        ropt = torch.optim.Adam(ref.parameters(), lr=1e-2)
        ref(x_all).square().mean().backward()
        ropt.step()
        rflat = torch.cat([p.detach().reshape(-1) for p in ref.parameters()])
        results.put((torch.equal(gathered[0], gathered[1]), (gathered[0] - rflat).abs().max().item(), hooks[0].bytes_last))
    # sharded render plumbing with a stand-in "model"
    model = lambda batch, **kw: {'rgb': batch['pixel_id'][:, 1:].float() * 2, 'depth': batch['pixel_id'][:, 0].float()}
    pid = torch.arange(30).reshape(10, 3)
    out = parallel.render_sharded(model, {'pixel_id': pid}, gather_keys=['rgb', 'depth'])
    assert torch.equal(out['rgb'], pid[:, 1:].float() * 2) and torch.equal(out['depth'], pid[:, 0].float())
    # the drop-in's own band sharding (models/*91.py render()): band -> per-ray dict -> one all-gather of a flat record
    band = parallel.eval_band(11, {})
    assert band == parallel.shard_bounds(11, rank, world)
    rows = torch.arange(11)[band[0]:band[1]]
    local = {'rgb': torch.stack([rows, rows * 2, rows * 3], 1).float(), 'depth': rows.float() + 0.5,
             'extr': rows.float()[:, None, None].expand(-1, 2, 2).contiguous()}
    full = parallel.gather_ray_outputs(local, 11)
    allr = torch.arange(11).float()
    assert torch.equal(full['rgb'], torch.stack([allr, allr * 2, allr * 3], 1)) and torch.equal(full['depth'], allr + 0.5)
    assert full['extr'].shape == (11, 2, 2) and torch.equal(full['extr'][:, 1, 1], allr)
    assert parallel.eval_band(1, {}) is None and parallel.eval_band(11, {'shard_eval_rays': False}) is None
    # this rank's rows of the global CPU draw
    torch.manual_seed(3)
    mine = parallel.rows_of_global_draw(lambda n: torch.rand([n, 4]), 3, (rank, world, 8, __import__('numpy').array([1, 4, 6])), chunk=8)
    torch.manual_seed(3)
    assert torch.equal(mine, torch.rand([8, 4])[[1, 4, 6]])
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_and_sharded_render_world2():
    ctx = mp.get_context('spawn')
    results = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    same, err, nbytes = results.get()
    assert same, 'ranks diverged after the all-reduced step'
    assert err <= 1e-6, err                               # mean of equal-share rank means == global mean
    assert nbytes == 4 * (5 * 7 + 7 + 7 * 2 + 2)


def _preproc_worker(rank, world, port, results):
    """DataPreprocessor91.select_batch_indices on two gloo ranks (no kernels involved: index bookkeeping only)."""
    import numpy
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from simple_rf_b200.dropin import callers as C
    C.prepare()
    import Trainer10
    from data_preprocessors.DataPreprocessorFactory01 import get_data_preprocessor
    cfg = C.use_dropin(C.complete_configs(C.load_shipped_configs(1142), [0], seed=5))
    cfg['data_loader']['num_rays'] = 48
    cfg['data_loader']['sparse_depth']['num_rays'] = 16
    raw = C.synthetic_raw_data('llff', 3, resolution=(24, 32), sparse_points=40, seed=1)
    out = {}
    for mode in ('single', 'strong', 'weak'):
        if mode != 'single' and not dist.is_initialized():
            parallel.init_from_env(backend='gloo')
        cfg['data_loader']['rank_sharding'] = mode if mode != 'single' else 'strong'
        Trainer10.init_seeds(5)
        pre = get_data_preprocessor(cfg, mode='train', raw_data_dict=raw)
        picks = []
        for it in range(3):                          # 3 x 48 > 24*32*3 / ... crosses no wrap; sparse-depth block wraps (120 points)
            d = pre.select_batch_indices(it, None)
            picks.append((d['indices'].numpy().copy(), d['indices_mask_nerf'].numpy().copy(), d.get('srf_shard')))
        out[mode] = picks
    # union of the strong shards == the single-process batch, block proportions kept
    ok = True
    for it in range(3):
        mine = torch.from_numpy(out['strong'][it][0])
        both = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        single_idx, single_mask, _ = out['single'][it]
        n_img = int(single_mask.sum())
        img = numpy.concatenate([b.numpy()[:n_img // world] for b in both])
        sd = numpy.concatenate([b.numpy()[n_img // world:] for b in both])
        ok &= numpy.array_equal(img, single_idx[:n_img]) and numpy.array_equal(sd, single_idx[n_img:])
        rk, wd, n_all, rows = out['strong'][it][2]
        ok &= (rk, wd, n_all) == (rank, world, single_idx.shape[0]) and numpy.array_equal(single_idx[rows], out['strong'][it][0])
        ok &= int(out['strong'][it][1].sum()) == n_img // world
    # weak: rank r's t-th batch is the single process's (t * world + r)-th draw
    Trainer10.init_seeds(5)
    cfg['data_loader']['rank_sharding'] = 'strong'
    dist.barrier()
    dist.destroy_process_group()
    pre = get_data_preprocessor(cfg, mode='train', raw_data_dict=raw)
    draws = [pre.select_batch_indices(i, None)['indices'].numpy().copy() for i in range(3 * world)]
    for it in range(3):
        ok &= numpy.array_equal(out['weak'][it][0], draws[it * world + rank])
    results.put(bool(ok))


def test_preprocessor_rank_sharding_gloo():
    from simple_rf_b200.dropin import callers as C
    if not C.available():
        pytest.skip('upstream tree not present')
    ctx = mp.get_context('spawn')
    results = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_preproc_worker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert results.get() and results.get()
