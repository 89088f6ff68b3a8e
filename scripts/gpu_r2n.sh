#!/bin/bash
# product-form setup kernel (k_setup_pf): parity (whole suite), memcheck, kernel times vs the round-1 setup kernels
TAG=${1:-r2n}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu: exit $? : $(tail -1 $O/pytest_gpu.log)" > $O/summary.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 6 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_pre_factor_blocks or (test_matches_reference_golden and (c3_b64 or c4_small))" > $O/memcheck.log 2>&1
echo "memcheck: $(grep 'ERROR SUMMARY' $O/memcheck.log | tail -1)" >> $O/summary.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_pre_factor_blocks" > $O/racecheck.log 2>&1
echo "racecheck: $(grep 'RACECHECK SUMMARY' $O/racecheck.log | tail -1)" >> $O/summary.txt
for sp in 1 0; do
echo "== QPB200_SETUP_PF=$sp (throughput mode)" >> $O/summary.txt
for cfg in "128 100 100 0" "8192 100 100 0" "1024 50 50 10" "64 200 200 0"; do
  QPB200_SETUP_PF=$sp QPB_KT_TWO=1 timeout 120 python scripts/kernel_times.py $cfg >> $O/summary.txt 2>&1
done
done
timeout 300 python scripts/c4_times.py >> $O/summary.txt 2>&1
QPB_BENCH_E2E=0 timeout 600 python bench.py --steps 24 --warmup 5 > $O/bench_v.json 2> $O/bench.err
cat $O/summary.txt; tail -4 $O/pytest_gpu.log; cat $O/bench_v.json; tail -3 $O/bench.err; grep -B2 -A10 "Invalid\|Race reported" $O/memcheck.log $O/racecheck.log | head -40
