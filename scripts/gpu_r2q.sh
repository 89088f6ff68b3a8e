#!/bin/bash
TAG=${1:-r2q}
O=gpurun_out/$TAG
mkdir -p $O
QPB200_NT512=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_matches_reference_golden and (c2 or c3_b64 or c4_small or band)" > $O/t.log 2>&1; echo "golden with 512-thread resident forward: exit $? : $(tail -1 $O/t.log)" > $O/summary.txt
for v in 0 2; do
echo "== QPB200_NT512=$v (latency mode)" >> $O/summary.txt
for cfg in "128 100 100 0" "64 50 50 10" "128 60 100 0"; do
  QPB200_NT512=$v timeout 120 python scripts/kernel_times.py $cfg >> $O/summary.txt 2>&1
done
done
QPB_BENCH_INFLIGHT=8 QPB_BENCH_E2E=0 timeout 600 python bench.py --steps 24 --warmup 5 > $O/bench_i8.json 2> $O/bench.err
QPB_BENCH_INFLIGHT=6 QPB_BENCH_E2E=0 timeout 600 python bench.py --steps 24 --warmup 5 > $O/bench_i6.json 2>> $O/bench.err
cat $O/summary.txt; cat $O/bench_i8.json $O/bench_i6.json; tail -3 $O/bench.err
