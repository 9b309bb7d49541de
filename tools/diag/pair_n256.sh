for v in pair_n256 single_n256; do echo "== $v"
SIMPLE_RF_B200_LIB=$PWD/variants/lib_$v.so timeout 120 python -m pytest tests/test_gpu_nerf_mlp.py -q -k "forward_vs_oracle or row_independent" 2>&1 | tail -12
SIMPLE_RF_B200_LIB=$PWD/variants/lib_$v.so timeout 120 python tools/mlp_microbench.py 2>&1 | grep "R=32768"
done
