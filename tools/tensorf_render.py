"""Simple-TensoRF frame render (576x1024, 331x368x220 grid, the blocky ~14 % alpha mask of bench.py) for profiling: python tools/tensorf_render.py [frames]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simple_rf_b200 import synthetic, _lib
from simple_rf_b200.models.SimpleTensoRF91 import AlphaGridMask, SimpleTensoRF

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device('cuda', 0)
cfg = synthetic.tensorf_configs(num_voxels=300 ** 3, augmentations=False)
mc = synthetic.scene_model_configs('re10k', num_views=3)
torch.manual_seed(0)
model = SimpleTensoRF(cfg, mc).to(dev).eval()
t = model.coarse_model
for p_ in t.matrices_density:
    p_.data.mul_(6.0)
vol = synthetic.blocky_alpha_volume(190, 10, 0.10, 0.004, torch.Generator().manual_seed(1))       # the mask bench.py uses
t.alpha_mask = AlphaGridMask(vol, t.bounding_box.cpu()).to(dev)
h, w = mc['resolution']
pid = torch.from_numpy(synthetic.frame_pixel_ids(h, w, view=0)).to(dev)


def render():
    with torch.no_grad():
        return model({'pixel_id': pid, 'num_frames': 3})


render()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(frames):
    out = render()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / frames
print(f'{ms:.2f} ms / frame, {pid.shape[0] / ms / 1e3:.2f} Mrays/s, {pid.shape[0] * int(t.num_samples) / ms / 1e6:.2f} Gsamples/s, acc mean {out["acc_coarse"].mean().item():.3f}')

# per-entry-point CUDA-event times of one more frame
_lib.TIMING = []
render()
torch.cuda.synchronize()
per = {}
for name, a, b, _ in _lib.TIMING:
    per[name] = per.get(name, 0.0) + a.elapsed_time(b)
_lib.TIMING = None
print('  '.join(f'{k} {v:.2f}' for k, v in sorted(per.items(), key=lambda kv: -kv[1])), '(ms / frame, timed entry points)')
