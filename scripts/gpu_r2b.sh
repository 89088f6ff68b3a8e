#!/bin/bash
# Round-2 second GPU pass: isolated tests, memcheck of the failing band, tiny-path and C4 timings, bench at restored defaults.
TAG=${1:-r2b}
O=gpurun_out/$TAG
mkdir -p $O
bash scripts/gpu_tests_isolated.sh $O/tests > $O/tests_summary.txt 2>&1
for c in band_smem_eq band_setup_eq; do
timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_matches_reference_golden and $c]" > $O/memcheck_$c.log 2>&1
done
echo "== kernel times (defaults: one QP per SM, 8-row substitution)" > $O/kernel_times.log
for cfg in "128 100 100 0" "1024 100 100 0" "1024 50 50 10" "64 200 200 0" "4096 10 5 0" "4096 24 24 0" "4096 32 24 8" "512 10 5 0"; do
  timeout 120 python scripts/kernel_times.py $cfg >> $O/kernel_times.log 2>&1
done
echo "-- tiny path off (256-thread fast kernels)" >> $O/kernel_times.log
for cfg in "4096 10 5 0" "4096 24 24 0" "4096 32 24 8" "512 10 5 0"; do
  QPB200_LIB=$PWD/build/variants/lib_notiny.so timeout 120 python scripts/kernel_times.py $cfg >> $O/kernel_times.log 2>&1
done
timeout 300 python scripts/c4_times.py > $O/c4_times.log 2>&1
QPB200_TIMING_LIB=$PWD/build/timing/t_r2.so timeout 120 python scripts/phase_timing.py > $O/phase.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
cat $O/tests_summary.txt; cat $O/kernel_times.log; cat $O/c4_times.log; head -c 1500 $O/bench.json
grep -A12 "Invalid\|ERROR SUMMARY" $O/memcheck_band_smem_eq.log | head -60
