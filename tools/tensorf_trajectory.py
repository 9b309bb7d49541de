"""Simple-TensoRF test-trajectory rendering (rgb + depth), frames sharded round-robin across ranks (BASELINE.json
configs[3]).  Launch: python tools/tensorf_trajectory.py [frames_per_rank]   or
python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/tensorf_trajectory.py [frames_per_rank]
Device-timed, barrier + synchronize on both sides, max over ranks; rank 0 prints one JSON line."""
import json, sys
from pathlib import Path
import torch
import torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simple_rf_b200 import parallel, synthetic
from simple_rf_b200.models.SimpleTensoRF91 import AlphaGridMask, SimpleTensoRF

frames_per_rank = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rank, world = parallel.init_from_env()
dev = torch.device('cuda', torch.cuda.current_device())
cfg = synthetic.tensorf_configs(num_voxels=300 ** 3, augmentations=False)
mc = synthetic.scene_model_configs('re10k', num_views=3)
torch.manual_seed(0)
model = SimpleTensoRF(cfg, mc).to(dev).eval()
t = model.coarse_model
with torch.no_grad():
    for p_ in t.matrices_density:
        p_.mul_(6.0)
vol = (torch.rand(190, 190, 190, generator=torch.Generator().manual_seed(1)) < 0.05).float()
t.alpha_mask = AlphaGridMask(vol, t.bounding_box.cpu()).to(dev)
h, w = mc['resolution']
pid = torch.from_numpy(synthetic.frame_pixel_ids(h, w, view=0)).to(dev)
total_frames = frames_per_rank * world
k_view = torch.tensor(mc['intrinsics'][0], dtype=torch.float32, device=dev)


def render(frame):
    pose = torch.tensor(synthetic.trajectory_pose(mc, frame / total_frames), dtype=torch.float32, device=dev)
    batch = {'pixel_id': pid, 'num_frames': 3, 'common_data': {'processed_view_pose': pose, 'view_intrinsic': k_view}}
    with torch.no_grad():
        out = model(batch, mode='static_camera')
    return out['rgb_coarse'], out['depth_coarse']


render(rank)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
check = 0.0
for f in range(rank, total_frames, world):
    rgb, depth = render(f)
    check += float(rgb.mean()) + float(depth.mean()) * 0      # device -> host read of the frame statistics
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    sec = ms.item() * 1e-3
    print(json.dumps({'workload': 'simple_tensorf_trajectory_render', 'n_gpus': world, 'frames': total_frames, 'frame': [h, w],
                      'samples_per_ray': int(t.num_samples), 'ms_total_max_over_ranks': round(ms.item(), 2),
                      'frames_per_sec': round(total_frames / sec, 3), 'rays_per_sec': round(total_frames * h * w / sec, 1),
                      'scaling': 'weak', 'mean_rgb_rank0': round(check / frames_per_rank, 5)}))
if world > 1:
    dist.destroy_process_group()
