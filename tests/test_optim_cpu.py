"""CPU: the Adam oracle is pinned against torch.optim.Adam (what the reference's OptimizerFactory02 builds); host logic of
the fused optimiser wrapper that needs no GPU."""
import numpy as np
import torch

from oracle import adam as OA


def test_adam_oracle_matches_torch():
    torch.manual_seed(0)
    shapes = [(256, 63), (256,), (3, 128), (1, 16, 9, 7)]
    params = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    opt = torch.optim.Adam([{'params': params[:2], 'lr': 5e-4}, {'params': params[2:], 'lr': 2e-2}], betas=(0.9, 0.999))
    state = [(p.detach().numpy().copy(), np.zeros(p.shape, np.float32), np.zeros(p.shape, np.float32)) for p in params]
    lrs = [5e-4, 5e-4, 2e-2, 2e-2]
    for step in range(1, 8):
        grads = [torch.randn(s) * 10.0 ** float(torch.randint(-4, 2, ())) for s in shapes]
        for p, g in zip(params, grads):
            p.grad = g.clone()
        opt.step()
        for i, g in enumerate(grads):
            state[i] = OA.adam_step(state[i][0], g.numpy(), state[i][1], state[i][2], step, lrs[i])
            ref = params[i].detach().numpy()
            st = opt.state[params[i]]
            # <= a couple of ulps: torch's CPU kernels fuse some of the multiply-adds
            assert np.abs(state[i][0] - ref).max() <= 4e-7 * max(1.0, np.abs(ref).max()), (step, i)
            for mine, theirs in ((state[i][1], st['exp_avg'].numpy()), (state[i][2], st['exp_avg_sq'].numpy())):
                assert np.abs(mine - theirs).max() <= 1e-6 * np.abs(theirs).max(), (step, i)
        for i, p in enumerate(params):                 # keep the two trajectories from drifting apart
            state[i] = (p.detach().numpy().copy(), opt.state[p]['exp_avg'].numpy().copy(), opt.state[p]['exp_avg_sq'].numpy().copy())


def test_runs_and_support():
    from simple_rf_b200 import optim
    assert optim._runs([True, True, False, True], [3, 3, 3, 3]) == [(0, 2), (3, 4)]
    assert optim._runs([True, True, True], [3, 0, 0]) == [(0, 1), (1, 3)]
    assert optim._runs([False, False], [0, 0]) == []
    p = torch.nn.Parameter(torch.zeros(4))
    assert not optim.supports(torch.optim.Adam([p]))                       # CPU parameters: left to torch
    assert not optim.supports(torch.optim.SGD([p], lr=0.1))
