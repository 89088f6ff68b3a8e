#!/bin/bash
# A/B pass over kernel build variants: build/variants/*.so (real builds: kernel_times) and build/timing/*.so
# (-DQPB_TIMING builds: phase accounting), plus parity tests and kernel times of the in-tree product build.
# Usage: gpurun -- 'bash scripts/gpu_ab.sh <tag>'
TAG=${1:-ab}
O=gpurun_out/$TAG
mkdir -p $O
echo "== product build" > $O/summary.txt
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_product.log 2>&1; echo "product pytest exit $?" >> $O/summary.txt
tail -3 $O/pytest_product.log >> $O/summary.txt
for cfg in "128 100 100 0" "1024 50 50 10" "1024 100 100 0"; do timeout 120 python scripts/kernel_times.py $cfg >> $O/summary.txt 2>&1; done
for v in build/variants/*.so; do
  echo "== $v" >> $O/summary.txt
  if [ -n "$AB_TESTS" ]; then
    QPB200_LIB=$PWD/$v timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_$(basename $v .so).log 2>&1; echo "pytest exit $?" >> $O/summary.txt
  fi
  for cfg in "128 100 100 0" "1024 50 50 10"; do QPB200_LIB=$PWD/$v timeout 120 python scripts/kernel_times.py $cfg >> $O/summary.txt 2>&1; done
done
for v in build/timing/*.so; do
  QPB200_TIMING_LIB=$PWD/$v timeout 120 python scripts/phase_timing.py > $O/phase_$(basename $v .so).log 2>&1
done
if [ -n "$AB_BENCH" ]; then timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; fi
cat $O/summary.txt
if [ -n "$AB_BENCH2" ]; then
  QPB_BENCH_INFLIGHT=3 QPB_BENCH_E2E_INFLIGHT=4 timeout 600 python bench.py > $O/bench_i3e4.json 2> $O/bench_i3e4.err
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench*.json")):
    try:
        d=json.load(open(f)); print(f, "value %.0f (%.3f ms/step, serial %.3f) e2e %.0f windows %s" % (d["value"], d["ms_per_step"], d["config"]["serial_ms_per_step"], d["e2e"]["value"], d["e2e"]["windows_ms"]))
    except Exception as e: print(f, "unreadable", e)
PY
