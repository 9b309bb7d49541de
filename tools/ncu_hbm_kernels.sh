#!/bin/bash
# ncu counters of the HBM / L2-bound kernels (north_star: "achieved HBM GB/s for the gather and
# compositing kernels"): dram bytes, L2 bytes, duration per launch of every srf:: kernel of
#   (a) the compositing micro-benchmark at 2^20 x 64 and 2^20 x 192    (composite_fwd / _bwd; one shape per run)
#   (b) two Simple-TensoRF training iterations at 331x368x220          (tensorf_mask, vm_density_*, vm_color_features_*, composite, tv, adam)
#   (c) one Simple-TensoRF test frame                                  (tensorf_march, vm_color_features_fwd, mlp rows)
#   (d) two Simple-NeRF training iterations, 4096 rays                 (nerf_mlp_fwd with saved tiles, nerf_mlp_dgrad, nerf_mlp_wgrad)
# An optional second argument runs only the legs whose name starts with it.
# Run on the GPU box:  bash tools/ncu_hbm_kernels.sh gpurun_out/r2_ncu     then here:  python tools/ncu_table.py gpurun_out/r2_ncu
set -u
OUT=${1:-gpurun_out/ncu_hbm}
ONLY=${2:-}
mkdir -p "$OUT"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
# ncu matches the function name without its namespace: list this library's kernel families
K='^(composite_|vm_|tensorf_|nerf_mlp_|sample_pdf|raygen|stratified|adam|tv_|alpha_|resample_|pack_alpha|compact_|scan_blocks|threshold_mask|scatter_rows|gather_rows|ray_|march_|assemble_|patch_|frame_)'
run() {  # name, skip, count, command...
  local name=$1 skip=$2 count=$3; shift 3
  if [ -n "$ONLY" ] && [[ "$name" != "$ONLY"* ]]; then return; fi
  timeout 600 ncu --metrics $M --clock-control none -k regex:"$K" -s "$skip" -c "$count" --csv --log-file "$OUT/$name.csv" "$@" > "$OUT/$name.log" 2>&1
  echo "$name rc=$?"
}
run composite_S64 0 120 python tools/hbm_microbench.py --probe 64
run composite_S192 0 120 python tools/hbm_microbench.py --probe 192
run tensorf_train 0 400 python tools/tensorf_train_step.py 2
run tensorf_frame 0 200 python tools/tensorf_render.py 1
run nerf_train 0 400 python tools/train_step.py 2
