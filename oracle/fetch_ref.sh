#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Ships the UNMODIFIED reference to the GPU box.
#
# The reference (AuroraEchos/Sound-Event-Localization-and-Detection) is pure Python: nothing is compiled here, the
# "build" of oracle/_ref is a verbatim copy of the nine .py files the hot path and its callers live in plus the five
# config files, taken from where they lie under /root/reference.  oracle/_ref/ is git-ignored (no reference source
# enters the history) but NOT gpurun-ignored, so it travels with the repo snapshot and
#   * `bench.py --impl reference` times the real model.SELD_Model on the box's host cores (cpu_baseline.kind
#     "reference"),
#   * the `-m gpu` tests run the unmodified model.py on top of the drop-in layer modules.
# Usage: bash oracle/fetch_ref.sh [reference root, default /root/reference]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${1:-/root/reference}"
DST="$HERE/_ref"
if [ ! -f "$SRC/model.py" ]; then
  echo "fetch_ref: no reference tree at $SRC (nothing copied)"; exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/quaternion" "$DST/dual_quaternion" "$DST/config"
cp "$SRC"/model.py "$SRC"/train.py "$SRC"/utility_functions.py "$SRC"/Dcase21_metrics.py "$SRC"/metrics.py "$DST"/
cp "$SRC"/quaternion/quaternion_ops.py "$SRC"/quaternion/quaternion_layers.py "$DST"/quaternion/
cp "$SRC"/dual_quaternion/dual_quaternion_ops.py "$SRC"/dual_quaternion/dual_quaternion_layers.py "$DST"/dual_quaternion/
cp "$SRC"/config/*.txt "$DST"/config/
( cd "$SRC" && sha256sum model.py train.py utility_functions.py Dcase21_metrics.py metrics.py quaternion/*.py \
    dual_quaternion/*.py config/*.txt ) > "$DST/SHA256SUMS"
echo "fetch_ref: copied $(find "$DST" -type f | wc -l) files into $DST"
