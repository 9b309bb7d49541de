#!/bin/bash
# Run on the GPU box: parity test + microbench for every variants/lib_*.so
cd "$(dirname "$0")/.."
for out in variants/lib_*.so; do
  echo "== $out"
  SIMPLE_RF_B200_LIB=$PWD/$out timeout 120 python -m pytest tests/test_gpu_nerf_mlp.py -x -q 2>&1 | tail -1
  SIMPLE_RF_B200_LIB=$PWD/$out timeout 120 python tools/mlp_microbench.py 2>&1 | grep "R=32768"
done
