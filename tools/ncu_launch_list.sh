#!/bin/bash
# Launch list of the driver's bench command (B200_PROFILING.md: `--metrics gpu__time_duration.sum --clock-control none`) for the main line
# and for the whole bench, plus one `--set full` capture of the dominant kernel's coarse-pass launch (roofline.traffic of bench.py).
# Numbers printed under ncu are never bench values: the per-launch times are cold-cache and serialised, the kernel's SHARE is what is compared.
#   bash tools/ncu_launch_list.sh gpurun_out/<dir>
set -u
OUT=${1:-gpurun_out/ncu_launches}
mkdir -p "$OUT"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file "$OUT/launches_main.csv" \
    python bench.py --main-only --steps 2 --warmup 1 > "$OUT/bench_main_under_ncu.log" 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file "$OUT/launches_all.csv" \
    python bench.py --steps 2 --warmup 1 > "$OUT/bench_all_under_ncu.log" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_fwd -s 6 -c 1 -f -o "$OUT/mlp_coarse_full" \
    python bench.py --main-only --steps 1 --warmup 3 > "$OUT/bench_full_under_ncu.log" 2>&1
ncu -i "$OUT/mlp_coarse_full.ncu-rep" --page raw --csv > "$OUT/mlp_coarse_full_raw.csv" 2>/dev/null
ls -la "$OUT"
