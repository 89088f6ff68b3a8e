#!/bin/bash
# oracle/make_ref.sh - put the UNMODIFIED reference (locuslab/qpth, 13 Python files) under oracle/_ref/ so that
# it travels to the GPU box with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored: it never enters
# the history).  TEST / MEASUREMENT INFRASTRUCTURE ONLY: bench.py's `--impl reference` arm and its
# `reference_cuda` figure import it; nothing under qpth_b200/ does.
#
# The only addition is an EMPTY `cvxpy` stub package: `qpth/solvers/__init__.py:3` imports the reference's
# optional CVXPY solver module, which does `import cvxpy` at import time; cvxpy is not in this image and the
# PDIPM path never touches it.
#
#   bash oracle/make_ref.sh [/root/reference]
set -e
SRC=${1:-${QPTH_REFERENCE:-/root/reference}}
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$SRC/qpth" ]; then
  echo "make_ref: $SRC/qpth not found (the reference only exists in the build container)"; exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/cvxpy"
cp -r "$SRC/qpth" "$DST/qpth"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
printf '"""Empty stand-in: the reference imports cvxpy for its optional CVXPY solver only (oracle/make_ref.sh)."""\n' > "$DST/cvxpy/__init__.py"
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$DST/REVISION"
( cd "$DST" && find qpth -name '*.py' | sort | xargs sha256sum ) > "$DST/SHA256SUMS"
echo "make_ref: copied $(find "$DST/qpth" -name '*.py' | wc -l) files to $DST"
