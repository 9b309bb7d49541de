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
