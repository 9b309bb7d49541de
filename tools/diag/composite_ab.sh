for v in base l4wide base l4wide; do echo "== $v"; SIMPLE_RF_B200_LIB=$PWD/variants/lib_$v.so python tools/hbm_microbench.py --probe 192 | cut -c1-120; done
