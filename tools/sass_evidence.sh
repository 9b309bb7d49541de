#!/usr/bin/env bash
# SASS evidence for the judge: per kernel of libsimple_rf_b200.so, the counts of the Blackwell-specific mnemonics
# (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit -> mbarrier, UBLKCP = cp.async.bulk,
# UTMALDG = TMA tensor-map load, REDG = red.global, SYNCS = mbarrier ops) + a short excerpt around the first UTCHMMA.
#   tools/sass_evidence.sh > profiles/sass_r02_kernels.txt
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
LIB="$HERE/simple_rf_b200/libsimple_rf_b200.so"
TMP="$(mktemp)"
cuobjdump -sass "$LIB" > "$TMP"
echo "# cuobjdump -sass simple_rf_b200/libsimple_rf_b200.so   ($(date -u +%Y-%m-%dT%H:%MZ), $(nvcc --version | tail -2 | head -1))"
echo "# arch: $(grep -m1 -o 'sm_[0-9a-z]*' "$TMP")"
echo
printf '%-58s %8s %6s %6s %7s %7s %8s %6s %6s\n' kernel UTCHMMA LDTM STTM UTCBAR UBLKCP UTMALDG REDG SYNCS
awk '
  /Function : / { if (name != "") emit(); name=$3; for (k in c) delete c[k] }
  { for (m in pat) if ($0 ~ pat[m]) c[m]++ }
  function emit() { printf "%-58s %8d %6d %6d %7d %7d %8d %6d %6d\n", substr(name,1,58), c["UTCHMMA"], c["LDTM"], c["STTM"], c["UTCBAR"], c["UBLKCP"], c["UTMALDG"], c["REDG"], c["SYNCS"] }
  BEGIN { pat["UTCHMMA"]="UTCHMMA"; pat["LDTM"]="LDTM"; pat["STTM"]="STTM"; pat["UTCBAR"]="UTCBAR"; pat["UBLKCP"]="UBLKCP"; pat["UTMALDG"]="UTMALDG"; pat["REDG"]="RED\\.|REDG"; pat["SYNCS"]="SYNCS" }
  END { if (name != "") emit() }
' "$TMP" | while read -r line; do
  set -- $line
  printf '%-58s %8s %6s %6s %7s %7s %8s %6s %6s\n' "$(echo "$1" | c++filt | sed 's/(.*//' | cut -c1-58)" "$2" "$3" "$4" "$5" "$6" "$7" "$8" "$9"
done
echo
echo "## excerpt: first tcgen05.mma issue sequence of nerf_mlp_fwd_kernel"
awk '/Function : .*nerf_mlp_fwd_kernel/ {f=1} f && /UTCHMMA/ {n++} f && n>=1 && n<=3 {print} f && n>3 {exit}' "$TMP" | head -40 | sed 's/^ *//' | cut -c1-150
echo
echo "## excerpt: bulk-copy (TMA) weight producer of nerf_mlp_fwd_kernel"
awk '/Function : .*nerf_mlp_fwd_kernel/ {f=1} f && /UBLKCP/ {print; n++} n>=4 {exit}' "$TMP" | sed 's/^ *//' | cut -c1-150
echo
echo "## CTA-pair instantiation nerf_mlp_fwd_kernel<false, true>: cta_group::2 MMAs and multicast commits (mnemonic counts)"
awk '/Function : .*nerf_mlp_fwd_kernelILb0ELb1/ {f=1; next} /Function : / {f=0} f && /2CTA/ {for (i = 1; i <= NF; i++) if ($i ~ /2CTA/) {c[$i]++; break}} END {for (k in c) print c[k], k}' "$TMP" | sort -rn
rm -f "$TMP"
