"""Build libsimple_rf_b200.so in-tree with nvcc for sm_100a (no torch headers: the ABI is plain C)."""
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / 'csrc'
LIB = HERE / 'libsimple_rf_b200.so'
SOURCES = ['rays_sampling.cu', 'composite.cu', 'nerf_mlp.cu', 'nerf_mlp_wgrad.cu', 'nerf_mlp_dgrad.cu', 'nerf_mlp_input_grad.cu', 'tensorf.cu', 'tensorf_march.cu', 'tensorf_surgery.cu', 'tensorf_cp.cu', 'losses.cu', 'optim.cu', 'batch.cu', 'output.cu']
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def nvcc_path():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError('nvcc not found')


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES if (CSRC / s).exists()] + list(CSRC.glob('*.cuh')) + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = nvcc_path()
    objs = []
    procs = []
    build_dir = HERE / 'build'
    build_dir.mkdir(exist_ok=True)
    for s in SOURCES:
        src = CSRC / s
        if not src.exists():
            continue
        obj = build_dir / (src.stem + '.o')
        cmd = [nvcc, *FLAGS, '-c', str(src), '-o', str(obj)]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
        if out.strip() and (verbose or p.returncode != 0):
            print(f'--- {s}\n{out}', file=sys.stderr)
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [nvcc, '-shared', '-o', str(LIB), *map(str, objs), '-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
