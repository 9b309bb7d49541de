"""Ray-sharded multi-GPU Simple-NeRF training iterations (torchrun): each rank takes its slice of the seeded 4096-ray batch,
one flat-bucket NCCL all-reduce per step.  Prints it/s (max over ranks, device-timed) and checks that ranks stay in lock-step."""
import os, sys
from pathlib import Path
import torch
import torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simple_rf_b200 import parallel, synthetic
from simple_rf_b200.models.SimpleNeRF91 import SimpleNeRF

rank, world = parallel.init_from_env()
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
torch.cuda.set_device(dev)
cfg = synthetic.nerf_configs(rng_mode='device')
mc = synthetic.scene_model_configs('llff', num_views=3)
torch.manual_seed(0)                                   # replicated parameters
model = SimpleNeRF(cfg, mc).to(dev).train()
opt = torch.optim.Adam(model.get_trainable_parameters(cfg['optimizers'][0]), betas=(0.9, 0.999))
model.optimizers = {'optimizer_main': opt}             # what Trainer10.py:61-62 does; attaches the gradient all-reduce
g = torch.Generator().manual_seed(2)
pid_all = torch.stack([torch.randint(0, 3, (4096,), generator=g), torch.randint(0, 1008, (4096,), generator=g),
                       torch.randint(0, 756, (4096,), generator=g)], 1).int()
target_all = torch.rand(4096, 3, generator=g)
mask_nerf = torch.zeros(4096, dtype=torch.bool); mask_nerf[:2048] = True
rows = parallel.shard_batch(mask_nerf, rank, world)
pid, target = pid_all[rows].to(dev), target_all[rows].to(dev)


def step():
    opt.zero_grad(set_to_none=True)
    o = model({'pixel_id': pid, 'num_frames': 3, 'iter_num': 0, 'sub_batch_index': 0})
    loss = sum(((o[k] - target) ** 2).mean() for k in ('rgb_coarse', 'rgb_fine', 'points_augmentation_rgb_coarse', 'views_augmentation_rgb_coarse'))
    loss.backward()
    opt.step()


for _ in range(3):
    step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 10
e0.record()
for _ in range(iters):
    step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
flat = torch.cat([p.detach().reshape(-1) for p in model.coarse_model.parameters()])
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([float(torch.equal(ref, flat))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
else:
    same = torch.ones(1)
if rank == 0:
    hooks = getattr(model, '_grad_allreduce', [])
    fused = getattr(model, '_fused_adam', [])
    nbytes = fused[0].bytes_reduced_last if fused else (hooks[0].bytes_last if hooks else 0)
    print(f'world {world}: {ms.item():.3f} ms / iteration ({1e3 / ms.item():.1f} it/s over a 4096-ray global batch), '
          f'ranks in lock-step: {bool(same.item())}, all-reduce bytes/step: {nbytes} '
          f"({'fused flat Adam' if fused else 'torch Adam + hook'})")
if world > 1:
    dist.destroy_process_group()
