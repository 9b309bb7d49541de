"""One Simple-TensoRF training iteration loop (4096 rays, main VM tensor + points-augmentation tensor, alpha mask set)
for profiling: python tools/tensorf_train_step.py [iters] [voxels_per_axis]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simple_rf_b200 import synthetic, _lib
from simple_rf_b200.models.SimpleTensoRF91 import AlphaGridMask, SimpleTensoRF

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
side = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device('cuda', 0)
cfg = synthetic.tensorf_configs(num_voxels=side ** 3, augmentations=True, rng_mode='device')
mc = synthetic.scene_model_configs('re10k', num_views=3)
torch.manual_seed(0)
model = SimpleTensoRF(cfg, mc).to(dev).train()
t = model.coarse_model
with torch.no_grad():
    for p_ in t.matrices_density:
        p_.mul_(6.0)
vol = (torch.rand(190, 190, 190, generator=torch.Generator().manual_seed(1)) < 0.05).float()
t.alpha_mask = AlphaGridMask(vol, t.bounding_box.cpu()).to(dev)
groups = model.get_trainable_parameters(cfg['optimizers'][0]) if hasattr(model, 'get_trainable_parameters') else None
opt = torch.optim.Adam(groups, betas=(0.9, 0.99))
model.optimizers = {'optimizer_nerf': opt}        # what Trainer10.py:59-62 does: the drop-in attaches its fused Adam step
h, w = mc['resolution']
g = torch.Generator().manual_seed(2)
pid = torch.stack([torch.randint(0, 3, (4096,), generator=g), torch.randint(0, w, (4096,), generator=g),
                   torch.randint(0, h, (4096,), generator=g)], 1).int().to(dev)
target = torch.rand(4096, 3, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    o = model({'pixel_id': pid, 'num_frames': 3, 'iter_num': 0, 'sub_batch_index': 0})
    loss = sum(((o[k] - target) ** 2).mean() for k in ('rgb_coarse', 'points_augmentation_rgb_coarse'))
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
print(f'{ms:.3f} ms / iteration ({1e3 / ms:.1f} it/s), S={int(t.num_samples)}, loss {loss.item():.4f}')
agg = {}
for name, a, b, work in _lib.TIMING:
    tt, ww = agg.get(name, (0.0, 0))
    agg[name] = (tt + a.elapsed_time(b), ww + 1)
for name, (tt, ww) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'  {name:28s} {tt / iters:8.3f} ms/iter  {ww / iters:5.1f} calls/iter')
