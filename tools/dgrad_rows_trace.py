"""Timeline of one 128-row tile of the colour (rows-mode) MLP inside the data-gradient kernel, plus device times of the three
rows-mode kernels (built with -DSRF_MLP_TRACE=1; tuning aid)."""
import ctypes, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from simple_rf_b200 import build as B
out = ROOT / 'gpurun_out' / 'lib_trace.so'
out.parent.mkdir(exist_ok=True)
src = [str(B.CSRC / f) for f in B.SOURCES]
subprocess.run([B.nvcc_path(), *B.FLAGS, '-DSRF_MLP_TRACE=1', '-shared', '-o', str(out)] + src, check=True)
os.environ['SIMPLE_RF_B200_LIB'] = str(out)
import torch
from simple_rf_b200 import _lib
from simple_rf_b200.nerf_program import PackedRowsMLP
dev = 'cuda'
g = torch.Generator().manual_seed(0)
m = PackedRowsMLP(72, 27, 3, prefix='mlp')
lin = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5).to(dev)
params = {'mlp.0.weight': lin(128, 30), 'mlp.0.bias': lin(128, 1)[:, 0], 'mlp.2.weight': lin(128, 128), 'mlp.2.bias': lin(128, 1)[:, 0],
          'mlp.4.weight': lin(3, 128), 'mlp.4.bias': torch.zeros(3, device=dev)}
m.refresh(params, lin(27, 72))
n = 148 * 128 * 40
rows = (torch.randn(n, 80, device=dev) * 0.1).to(torch.bfloat16)
count = torch.tensor([n], dtype=torch.int32, device=dev)
g_rgb = torch.randn(n, 3, device=dev)
for _ in range(2):
    rgb, acts = m.forward(rows, count, n, save=True)
    m.backward(acts, rgb, g_rgb, n, count=count)
torch.cuda.synchronize()
_lib.TIMING = []
rgb, acts = m.forward(rows, count, n, save=True)
m.backward(acts, rgb, g_rgb, n, count=count)
torch.cuda.synchronize()
for name, e0, e1, work in _lib.TIMING:
    ms = e0.elapsed_time(e1)
    print(f'{name}: {ms:.3f} ms for {n} rows = {ms * 1e-3 * 1.9e9 / 40:.0f} cycles per tile per SM (at 1.9 GHz)')
_lib.TIMING = None
buf = (ctypes.c_longlong * 2048)()
lib = _lib.load()
lib.srf_debug_dgrad_trace.argtypes = [ctypes.c_void_p]
assert lib.srf_debug_dgrad_trace(buf) == 0
t = list(buf)
prog = m.backward_plan.program
t0 = t[1024]
for l in range(prog.num_layers):
    L = prog.layers[l]
    mma = [f'kb{kb}: A@{t[16 + (l * 4 + kb) * 4] - t0} W@{t[16 + (l * 4 + kb) * 4 + 1] - t0} issued@{t[16 + (l * 4 + kb) * 4 + 2] - t0}' for kb in range(L.num_kblocks)]
    nb = L.n_out // 64
    epi = ' '.join(f'[pub@{t[1024 + l * 16 + 1 + 2 * kb] - t0} ship@{t[1024 + l * 16 + 2 + 2 * kb] - t0}]' for kb in range(nb))
    print(f'layer {l} (n_out {L.n_out}, dz_slot {L.dz_slot}, rows_cols {L.rows_cols}): MMA ' + ' | '.join(mma))
    print(f'        EPI d_full@{t[1024 + l * 16] - t0} {epi}')
