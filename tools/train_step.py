"""One Simple-NeRF training iteration loop (4096 rays, main + augmented MLPs) for profiling: python tools/train_step.py [iters]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simple_rf_b200 import synthetic, _lib
from simple_rf_b200.models.SimpleNeRF91 import SimpleNeRF

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device('cuda', 0)
cfg = synthetic.nerf_configs(rng_mode='device')
mc = synthetic.scene_model_configs('llff', num_views=3)
torch.manual_seed(0)
model = SimpleNeRF(cfg, mc).to(dev).train()
opt = torch.optim.Adam(model.get_trainable_parameters(cfg['optimizers'][0]), betas=(0.9, 0.999))
model.optimizers = {'optimizer_nerf': opt}        # what Trainer10.py:59-62 does: the drop-in attaches its fused Adam step
g = torch.Generator().manual_seed(2)
pid = torch.stack([torch.randint(0, 3, (4096,), generator=g), torch.randint(0, 1008, (4096,), generator=g),
                   torch.randint(0, 756, (4096,), generator=g)], 1).int().to(dev)
target = torch.rand(4096, 3, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    o = model({'pixel_id': pid, 'num_frames': 3, 'iter_num': 0, 'sub_batch_index': 0})
    loss = sum(((o[k] - target) ** 2).mean() for k in ('rgb_coarse', 'rgb_fine', 'points_augmentation_rgb_coarse', 'views_augmentation_rgb_coarse'))
    loss = loss + 0.1 * (o['depth_coarse'] - o['points_augmentation_depth_coarse'].detach()).square().mean()
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
_lib.TIMING = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f'{ms:.3f} ms / iteration ({1e3 / ms:.1f} it/s), loss {loss.item():.4f}')
agg = {}
for name, a, b, work in _lib.TIMING:
    t, w = agg.get(name, (0.0, 0.0))
    agg[name] = (t + a.elapsed_time(b), w + work)
for name, (t, w) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'  {name:24s} {t / iters:8.3f} ms/iter   {w / (t * 1e-3) / 1e12:8.1f} TFLOP/s (algorithmic)')
