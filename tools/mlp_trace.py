"""Timeline of one 128-row tile inside the fused MLP kernel (build with -DSRF_MLP_TRACE=1; tuning aid)."""
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
from simple_rf_b200 import _lib, nerf_program
from oracle import nerf_mlp as M
cfg = json.loads((ROOT / 'tests/golden/nerf_configs.json').read_text())['configs']['model']['coarse_model']
dev = 'cuda'
params = {k: v.to(dev) for k, v in M.init_mlp_params(cfg, torch.Generator().manual_seed(0)).items()}
packed = nerf_program.PackedMLP(cfg).refresh(params)
R, S = 32768, 192
o = torch.rand(R, 3, device=dev) - .5; d = torch.rand(R, 3, device=dev) - .5
vd = torch.nn.functional.normalize(torch.randn(R, 3, device=dev), dim=-1); z = torch.rand(R, S, device=dev)
for _ in range(3): packed.forward(o, d, z, vd)
buf = (ctypes.c_longlong * 4096)()
lib = _lib.load()
lib.srf_debug_mlp_trace.argtypes = [ctypes.c_void_p]
assert lib.srf_debug_mlp_trace(buf) == 0
t = list(buf)
prog = packed.program
t0 = t[1024]                      # epilogue sees layer 0 complete
print('encoding warps (run ahead): E published', t[1] - t[0], 'cycles after region 0 was free; relative to t0:', t[0] - t0, t[1] - t0)
s = 0
SPLIT = int(os.environ.get('SRF_MLP_SPLIT', '1'))
for l in range(prog.num_layers):
    L = prog.layers[l]
    mma = []
    halves = (L.n // 128) if SPLIT else 1
    for h in range(halves):
        for kb in range(L.num_kblocks):
            a, w, i = (t[16 + s * 4 + k] - t0 for k in range(3))
            mma.append(f'h{h}kb{kb}: A@{a} W@{w} issued@{i}')
            s += 1
    e = [t[1024 + l * 16 + k] - t0 for k in range(10)]
    print(f'layer {l}: MMA ' + ' | '.join(mma))
    print(f'         EPI d_full@{e[0]} ' + ' '.join(f'[ld@{e[1 + 2 * k]} st@{e[2 + 2 * k]}]' for k in range(L.n // 64)) + f' d_full(h1)@{e[9]}')
