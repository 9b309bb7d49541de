cd /root/repo
mkdir -p gpurun_out/r3c
timeout 200 python -m pytest tests/test_gpu_composite.py -x -q 2>&1 | tail -1
for v in b43 b44 b53 b54; do
  SIMPLE_RF_B200_LIB=$PWD/variants/lib_$v.so timeout 200 python tools/hbm_microbench.py --sweep > gpurun_out/r3c/sweep_$v.jsonl 2>/dev/null
  echo "== $v"; python tools/sweep_table.py gpurun_out/r3c/sweep_$v.jsonl | sed -n 10,18p
done
python tools/sweep_table.py gpurun_out/r3c/sweep_b54.jsonl | sed -n 1,9p
