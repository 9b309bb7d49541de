#!/bin/bash
# One `ncu --set full` capture of the top kernels (B200_PROFILING.md): the fused MLP forward as one CTA per tile and as CTA pairs
# (32 768 rays x 192 samples, main variant), and the TensoRF march v2 (one 65 536-ray launch group of the bench scene).
#   bash tools/ncu_full_top_kernels.sh gpurun_out/r2_full     then here:  python tools/ncu_full_summary.py gpurun_out/r2_full
set -u
OUT=${1:-gpurun_out/ncu_full}
mkdir -p "$OUT"
SRF_MLP_PAIR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^nerf_mlp_fwd_kernel' -s 28 -c 1 -f -o "$OUT/mlp_fwd_single" python tools/mlp_microbench.py > "$OUT/mlp_fwd_single.log" 2>&1; echo "single rc=$?"
SRF_MLP_PAIR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^nerf_mlp_fwd_kernel' -s 28 -c 1 -f -o "$OUT/mlp_fwd_pair" python tools/mlp_microbench.py > "$OUT/mlp_fwd_pair.log" 2>&1; echo "pair rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^tensorf_march_kernel' -s 9 -c 1 -f -o "$OUT/march" python tools/tensorf_render.py 1 > "$OUT/march.log" 2>&1; echo "march rc=$?"
ls -la "$OUT"
