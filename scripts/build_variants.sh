#!/bin/bash
# Build A/B variants of the kernels for scripts/gpu_ab.sh / gpu_probe.sh (they travel to the GPU box under build/).
#   scripts/build_variants.sh real  "NAME:-DFLAG=1 -DOTHER=0" ...   -> build/variants/lib_NAME.so
#   scripts/build_variants.sh timing "NAME:-DFLAG=1" ...            -> build/timing/t_NAME.so   (adds -DQPB_TIMING)
# Flags: QPB_PFORM, QPB_VECWARP, QPB_VG_SMEM, QPB_CHAIN_V2, QPB_TIMING_REPEAT, QPB_TIMING_PROBES (see profiles/r1_experiments.md).
# Every variant is the full three-unit library (qpth_b200/build.py): the flags reach all units.
set -e
cd "$(dirname "$0")/.."
kind=$1; shift
if [ "$kind" = timing ]; then dir=build/timing; pre=t_; extra=-DQPB_TIMING; else dir=build/variants; pre=lib_; extra=; fi
mkdir -p $dir
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( python -c "import sys; from qpth_b200 import build; build.build(force=True, extra=sys.argv[2:], out=sys.argv[1])" $PWD/$dir/$pre$name.so $extra $flags 2>&1 | grep -E "error" || true; echo "built $dir/$pre$name.so [$flags]" ) &
done
wait
