#!/bin/bash
# Build the schedule variants of the fused MLP kernel and time each (run on the GPU box; nvcc is there too).
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/variants
for g in 2 4; do for pf in 0; do
  out=gpurun_out/variants/lib_g${g}_pf${pf}.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    --expt-relaxed-constexpr -DSRF_MLP_GROUPS=$g -DSRF_MLP_PREFETCH=$pf -shared -o $out \
    simple_rf_b200/csrc/rays_sampling.cu simple_rf_b200/csrc/composite.cu simple_rf_b200/csrc/nerf_mlp.cu simple_rf_b200/csrc/tensorf.cu &
done; done; wait
for g in 2 4; do for pf in 0; do
  echo "== GROUPS=$g PREFETCH=$pf"
  SIMPLE_RF_B200_LIB=$PWD/gpurun_out/variants/lib_g${g}_pf${pf}.so timeout 120 python -m pytest tests/test_gpu_nerf_mlp.py -x -q 2>&1 | tail -1
  SIMPLE_RF_B200_LIB=$PWD/gpurun_out/variants/lib_g${g}_pf${pf}.so timeout 120 python tools/mlp_microbench.py 2>&1 | grep "R=32768"
done; done
