"""Quick device-timed throughput of the fused MLP kernel (not the contract bench; used while tuning)."""
import json, sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from simple_rf_b200 import nerf_program
from oracle import nerf_mlp as M

cfgs = json.loads((Path(__file__).resolve().parent.parent / 'tests/golden/nerf_configs.json').read_text())['configs']['model']
variants = {'main': cfgs['coarse_model'], 'pa': cfgs['augmentations'][0]['coarse_model'], 'va': cfgs['augmentations'][1]['coarse_model']}
dev = 'cuda'
g = torch.Generator().manual_seed(0)
for name, cfg in variants.items():
    params = {k: v.to(dev) for k, v in M.init_mlp_params(cfg, g).items()}
    packed = nerf_program.PackedMLP(cfg).refresh(params)
    for R, S in ((4096, 64), (4096, 192), (32768, 192)):
        o = torch.rand(R, 3, device=dev) - .5; d = torch.rand(R, 3, device=dev) - .5
        vd = torch.nn.functional.normalize(torch.randn(R, 3, device=dev), dim=-1); z = torch.rand(R, S, device=dev)
        for _ in range(3): packed.forward(o, d, z, vd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n): packed.forward(o, d, z, vd)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2 * packed.macs_per_sample * R * S
        print(f'{name} R={R} S={S}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {R * S / ms / 1e3:.1f} Msamples/s')
