"""Timeline of one 128-row tile inside the fused data-gradient kernel (built with -DSRF_MLP_TRACE=1; tuning aid)."""
import ctypes, json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
out = ROOT / 'gpurun_out' / 'lib_trace.so'
out.parent.mkdir(exist_ok=True)
from simple_rf_b200 import build as B
src = [str(B.CSRC / f) for f in B.SOURCES]
subprocess.run([B.nvcc_path(), *B.FLAGS, '-DSRF_MLP_TRACE=1', '-shared', '-o', str(out)] + src, check=True)
os.environ['SIMPLE_RF_B200_LIB'] = str(out)
import torch
from simple_rf_b200 import _lib, nerf_program as NP
from oracle import nerf_mlp as M
cfg = json.loads((ROOT / 'tests/golden/nerf_configs.json').read_text())['configs']['model']['coarse_model']
dev = 'cuda'
params = {k: v.to(dev) for k, v in M.init_mlp_params(cfg, torch.Generator().manual_seed(0)).items()}
packed = NP.PackedMLP(cfg).refresh(params)
R, S = 4096, 192
o = torch.rand(R, 3, device=dev) - .5; d = torch.rand(R, 3, device=dev) - .5
vd = torch.nn.functional.normalize(torch.randn(R, 3, device=dev), dim=-1); z = torch.rand(R, S, device=dev)
gs = torch.randn(R, S, 1, device=dev); gc = torch.randn(R, S, 3, device=dev)
for _ in range(3):
    sigma, rgb, acts = packed.forward(o, d, z, vd, save=True)
    NP.mlp_backward(packed, packed.flat, acts, sigma, rgb, gs, gc)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 2048)()
lib = _lib.load()
lib.srf_debug_dgrad_trace.argtypes = [ctypes.c_void_p]
assert lib.srf_debug_dgrad_trace(buf) == 0
t = list(buf)
prog = packed.backward_plan.program
t0 = t[1024]
for l in range(prog.num_layers):
    L = prog.layers[l]
    mma = [f'kb{kb}: A@{t[16 + (l * 4 + kb) * 4] - t0} W@{t[16 + (l * 4 + kb) * 4 + 1] - t0} issued@{t[16 + (l * 4 + kb) * 4 + 2] - t0}' for kb in range(L.num_kblocks)]
    nb = L.n_out // 64
    epi = ' '.join(f'[pub@{t[1024 + l * 16 + 1 + 2 * kb] - t0} ship@{t[1024 + l * 16 + 2 + 2 * kb] - t0}]' for kb in range(nb))
    print(f'layer {l} (n_out {L.n_out}, dz_slot {L.dz_slot}): MMA ' + ' | '.join(mma))
    print(f'        EPI d_full@{t[1024 + l * 16] - t0} {epi}')
