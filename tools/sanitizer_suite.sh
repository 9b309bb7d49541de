#!/bin/bash
# compute-sanitizer over the small-shape GPU parity tests (VERDICT r1 missing #7).  memcheck on every kernel family,
# racecheck on the shared-memory heavy ones.  The tcgen05 kernels run under memcheck too (bounded by `timeout`: the tool
# serialises warps and the mbarrier pipelines are slow under it).
#   bash tools/sanitizer_suite.sh gpurun_out/r2_sanitizer [only-suites-whose-name-starts-with]
set -u
OUT=${1:-gpurun_out/sanitizer}
ONLY=${2:-}
mkdir -p "$OUT"
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, tool, timeout, pytest args...
  local name=$1 tool=$2 limit=$3; shift 3
  if [ -n "$ONLY" ] && [[ "$name" != "$ONLY"* ]]; then return; fi
  timeout "$limit" $SAN --tool "$tool" --print-limit 5 --error-exitcode 99 python -m pytest -q -x -p no:cacheprovider "$@" > "$OUT/$name.$tool.log" 2>&1
  local rc=$?
  echo "$name $tool rc=$rc $(grep -E 'ERROR SUMMARY|passed|failed' "$OUT/$name.$tool.log" | tr '\n' ' ')"
}
run sampling memcheck 240 tests/test_gpu_sampling.py
run composite memcheck 240 tests/test_gpu_composite.py
run tensorf memcheck 300 tests/test_gpu_tensorf.py -k "mask or density or color or compaction"
run surgery memcheck 240 tests/test_gpu_surgery.py -k "not full_size and not 48-368"
run batch_losses_optim memcheck 240 tests/test_gpu_batch.py tests/test_gpu_losses.py tests/test_gpu_optim.py
run composite racecheck 240 tests/test_gpu_composite.py -k "golden"
run sampling racecheck 240 tests/test_gpu_sampling.py
run tensorf racecheck 240 tests/test_gpu_tensorf.py -k "mask or density"
run mlp memcheck 240 tests/test_gpu_nerf_mlp.py -k "37-64 or 41-64 or 1-1"
run mlp_bwd memcheck 300 tests/test_gpu_nerf_mlp_bwd.py
run nerf_model memcheck 300 tests/test_gpu_nerf_model.py -k "golden or split or retraw or gradients"
run tensorf_model memcheck 300 tests/test_gpu_tensorf.py -k "dropin or golden"
run composite_all racecheck 240 tests/test_gpu_composite.py
run surgery racecheck 240 tests/test_gpu_surgery.py -k "not full_size and not 48-368"
run mlp racecheck 240 tests/test_gpu_nerf_mlp.py -k "37-64 or 41-64 or 1-1"
run cp_tensor memcheck 300 tests/test_gpu_tensorf_cp.py
run cp_tensor racecheck 300 tests/test_gpu_tensorf_cp.py -k "density or color_rows or schedule"
run camera_grad memcheck 300 tests/test_gpu_camera_grad.py tests/test_gpu_tensorf.py -k "input_gradients or struct or color_rows_and_mlp"
run camera_grad racecheck 300 tests/test_gpu_camera_grad.py -k "input_gradients and main"
