"""TEST INFRASTRUCTURE (CPU oracle): restatement of the Adam step the reference's optimisers take
(src/optimizers/OptimizerFactory02.py:9-22 builds torch.optim.Adam with lr / betas only; src/Trainer10.py:109-110 calls
step()).  The arithmetic lives in PyTorch (torch/optim/adam.py::_single_tensor_adam, pinned version in SURVEY.md §8c); this is
its numpy fp32 restatement, pinned against torch.optim.Adam itself in tests/test_optim_cpu.py (<= 2 ulp per step)."""
import numpy as np


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """One step on fp32 arrays (returns new p, m, v); `step` counts from 1.  Scalars are Python doubles, as in torch."""
    f = np.float32
    p, g, m, v = (np.asarray(x, dtype=f) for x in (p, g, m, v))
    if weight_decay != 0.0:
        g = g + f(weight_decay) * p
    m = m + (g - m) * f(1.0 - beta1)                                   # exp_avg.lerp_(grad, 1 - beta1)
    v = v * f(beta2) + (f(1.0 - beta2) * g) * g                        # exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) * f(1.0 / np.sqrt(bc2)) + f(eps)                # (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = p - f(step_size) * (m / denom)                                 # param.addcdiv_(exp_avg, denom, value=-step_size)
    return p.astype(f), m.astype(f), v.astype(f)
