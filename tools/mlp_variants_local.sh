#!/bin/bash
# Build MLP-kernel variants HERE (cross-compile) into variants/ (travels with gpurun; *.so is git-ignored).
# usage: tools/mlp_variants_local.sh "name -DSRF_MLP_X=..." ...
cd "$(dirname "$0")/.."
mkdir -p variants
SRC=$(python -c "from simple_rf_b200 import build as B; print(' '.join(str(B.CSRC / s) for s in B.SOURCES))")
for v in "$@"; do set -- $v; name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    --expt-relaxed-constexpr "$@" -shared -o variants/lib_$name.so $SRC &
done; wait; ls -la variants
