# CTA-pair MLP forward vs one CTA per tile on one box: parity tests, then the micro-benchmark
echo "== pair parity"; SRF_MLP_PAIR=1 timeout 90 python -m pytest tests/test_gpu_nerf_mlp.py -x -q -k "forward_vs_oracle or row_independent" 2>&1 | tail -3
for v in 0 1 0 1; do echo "== SRF_MLP_PAIR=$v"; SRF_MLP_PAIR=$v timeout 120 python tools/mlp_microbench.py 2>&1 | grep "R=32768"; done
