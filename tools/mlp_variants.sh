#!/bin/bash
# Build the schedule variants of the fused MLP kernel and time each (run on the GPU box; nvcc is there too).
# usage: tools/mlp_variants.sh "2 0" "2 1" "4 0" ...   (GROUPS PREFETCH pairs)
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/variants
SRC=$(python -c "from simple_rf_b200 import build as B; print(' '.join(str(B.CSRC / s) for s in B.SOURCES))")
[ $# -eq 0 ] && set -- "2 0" "2 1" "4 0" "4 1"
for v in "$@"; do set -- $v; g=$1; pf=$2
  out=gpurun_out/variants/lib_g${g}_pf${pf}.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    --expt-relaxed-constexpr -DSRF_MLP_GROUPS=$g -DSRF_MLP_PREFETCH=$pf -shared -o $out $SRC &
done; wait
for out in gpurun_out/variants/lib_g*_pf*.so; do
  echo "== $out"
  SIMPLE_RF_B200_LIB=$PWD/$out timeout 120 python -m pytest tests/test_gpu_nerf_mlp.py -x -q 2>&1 | tail -1
  SIMPLE_RF_B200_LIB=$PWD/$out timeout 120 python tools/mlp_microbench.py 2>&1 | grep "R=32768"
done
