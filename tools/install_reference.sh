#!/usr/bin/env bash
# Install the UNMODIFIED upstream Simple-RF tree next to this repo so that it travels to the GPU box with `gpurun`:
#   baseline/_ref/src                      <- /root/reference/src            (pure Python, no setup.py: `pip install` has nothing to build)
#   baseline/_ref/runs/training/train{1061,1142,0212}   <- the shipped Configs.json / ModelConfigs.json fixtures
# baseline/_ref/ is git-ignored (never part of the history) but NOT gpurun-ignored.  The tree is used by
#   * bench.py --impl reference   (reference CPU arm: the reference's own classes through Trainer10 / Tester07),
#   * tests/test_gpu_reference_callers.py (unmodified Trainer.train_one_iter / NerfTester.predict_frame driving the drop-in
#     classes, compared with the reference's own models on the same box),
# and by nothing under simple_rf_b200/ that the product path needs.
set -euo pipefail
SRC="${SIMPLE_RF_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$HERE/baseline/_ref"
if [ ! -f "$SRC/src/models/ModelFactory02.py" ]; then
  echo "install_reference: no upstream checkout at $SRC (nothing to do)" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/runs/training"
cp -r "$SRC/src" "$DST/src"
find "$DST/src" -name '__pycache__' -type d -prune -exec rm -rf {} +
for run in train1061 train1142 train0212; do
  cp -r "$SRC/runs/training/$run" "$DST/runs/training/$run"
done
cp "$SRC/LICENSE" "$DST/LICENSE" 2>/dev/null || true
( cd "$SRC" && find src -type f -name '*.py' -print0 | sort -z | xargs -0 sha256sum ) > "$DST/SOURCE_SHA256SUMS"
echo "installed $(find "$DST/src" -name '*.py' | wc -l) reference modules into $DST"
