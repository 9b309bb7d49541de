"""Timeline of two consecutive 128-row tiles of the colour (rows-mode) MLP inside the fused MLP kernel
(build with -DSRF_MLP_TRACE=1; tuning aid)."""
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
for _ in range(3):
    m.forward(rows, count, n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.forward(rows, count, n); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f'{ms:.3f} ms for {n} rows: {n / ms / 1e6:.2f} G rows/s, {ms * 1e-3 * 1.9e9 / 40:.0f} cycles per tile per SM (at 1.9 GHz)')
buf = (ctypes.c_longlong * 4096)()
lib = _lib.load()
lib.srf_debug_mlp_trace.argtypes = [ctypes.c_void_p]
assert lib.srf_debug_mlp_trace(buf) == 0
t = list(buf)
prog = m.program
t0 = t[1024]
for k in (0, 1):
    o = k * 2048
    print(f'--- tile {2 + k} (cycles relative to tile 2 layer 0 d_full)')
    print('  enc: region-0 free @', t[o + 0] - t0, ' E published @', t[o + 1] - t0)
    s = 0
    for l in range(prog.num_layers):
        L = prog.layers[l]
        mma = []
        for kb in range(L.num_kblocks):
            a, w, i = (t[o + 16 + s * 4 + q] - t0 for q in range(3))
            mma.append(f'kb{kb}: A@{a} W@{w} issued@{i}')
            s += 1
        e = [t[o + 1024 + l * 16 + q] - t0 for q in range(1 + 2 * (L.n // 64))]
        print(f'  layer {l}: MMA ' + ' | '.join(mma))
        print(f'           EPI d_full@{e[0]} ' + ' '.join(f'[ld@{e[1 + 2 * q]} st@{e[2 + 2 * q]}]' for q in range(L.n // 64)),
              f' head: enter @{t[o + 2010 + l] - t0} after bar @{t[o + 2020 + l] - t0} done @{t[o + 2000 + l] - t0}' if L.head else '')
