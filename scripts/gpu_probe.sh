#!/bin/bash
# Timing-build probes (phase accounting of every build/timing/*.so) + kernel times (and parity, PROBE_TESTS=1) of
# build/variants/*.so. Usage: gpurun -- 'bash scripts/gpu_probe.sh <tag>'
TAG=${1:-probe}
O=gpurun_out/$TAG
mkdir -p $O
for v in build/timing/*.so; do
  QPB200_TIMING_LIB=$PWD/$v timeout 120 python scripts/phase_timing.py > $O/phase_$(basename $v .so).log 2>&1
done
for v in build/variants/*.so; do
  echo "== $v" >> $O/summary.txt
  QPB200_LIB=$PWD/$v timeout 120 python scripts/kernel_times.py 128 100 100 0 >> $O/summary.txt 2>&1
  if [ -n "$PROBE_TESTS" ]; then
    QPB200_LIB=$PWD/$v timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_$(basename $v .so).log 2>&1; echo "pytest exit $?" >> $O/summary.txt
    tail -2 $O/pytest_$(basename $v .so).log >> $O/summary.txt
  fi
done
cat $O/summary.txt
