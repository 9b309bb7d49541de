#!/bin/bash
# Same-box A/B of compositing build variants: every variants/lib_*.so (tools/mlp_variants_local.sh "name -DSRF_CMP_...=..") runs the
# parity tests and the BASELINE configs[4] sweep; tables under gpurun_out/<out>/.   bash tools/diag/composite_ab.sh gpurun_out/r3h
cd "$(dirname "$0")/../.."
OUT=${1:-gpurun_out/composite_ab}
mkdir -p "$OUT"
for lib in variants/lib_*.so; do
  v=$(basename "$lib" .so); v=${v#lib_}
  SIMPLE_RF_B200_LIB=$PWD/$lib timeout 100 python -m pytest tests/test_gpu_composite.py -x -q 2>&1 | tail -1
  SIMPLE_RF_B200_LIB=$PWD/$lib timeout 200 python tools/hbm_microbench.py --sweep > "$OUT/sweep_$v.jsonl" 2>/dev/null
  echo "== $v"; python tools/sweep_table.py "$OUT/sweep_$v.jsonl" | sed -n 4,18p
done
