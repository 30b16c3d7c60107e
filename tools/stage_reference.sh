#!/bin/bash
# Stage an UNMODIFIED copy of the reference's python package + its bundled cinderella sample under baseline/_ref
# (git-ignored, so it never enters history; NOT gpurun-ignored, so it travels to the GPU box with the snapshot).
# tests/test_e2e_cinderella.py uses it there to run the reference's own ComoRAG.py -- once on the reference classes
# (CPU) and once on the comorag_b200 shim (cuda:0) -- because /root/reference does not exist on the GPU box.
#   bash tools/stage_reference.sh [/root/reference]
set -euo pipefail
SRC="${1:-/root/reference}"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
DST="$ROOT/baseline/_ref"
[ -d "$SRC/src/comorag" ] || { echo "no reference tree at $SRC" >&2; exit 1; }
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/src" "$DST/src"
mkdir -p "$DST/dataset"
cp -r "$SRC/dataset/cinderella" "$DST/dataset/cinderella"
cp "$SRC/LICENSE" "$DST/LICENSE"
find "$DST" -name __pycache__ -prune -exec rm -rf {} +
chmod -R u+w "$DST"
( cd "$SRC" && find src dataset/cinderella -type f ! -name '*.pyc' -print0 | sort -z | xargs -0 sha256sum ) > "$DST/SHA256SUMS"
echo "staged $(find "$DST" -type f | wc -l) files under $DST"
