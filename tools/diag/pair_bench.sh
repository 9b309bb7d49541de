# main bench line (power-capped, 762 048-ray frames) with and without CTA pairs, same box
for v in 0 1 0 1; do echo "== SRF_MLP_PAIR=$v"; SRF_MLP_PAIR=$v python bench.py --main-only --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), 'rays/s', round(d['ms_per_step'], 2), 'ms', 'mlp', round(d['roofline']['achieved'], 1), 'TF frac', round(d['roofline']['frac'], 3), d['clocks'])"; done
echo "== pair N=256 variant"; SIMPLE_RF_B200_LIB=$PWD/variants/lib_pair_n256.so python bench.py --main-only --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), 'rays/s', round(d['ms_per_step'], 2), 'ms', 'mlp', round(d['roofline']['achieved'], 1), 'TF frac', round(d['roofline']['frac'], 3), d['clocks'])"
