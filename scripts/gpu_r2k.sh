#!/bin/bash
# 192-thread (three QPs per SM) and 512-thread (large orders) builds of the product-form kernels: parity + kernel times
TAG=${1:-r2k}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "product_form_kernels or families or two_per_sm" > $O/t_pf.log 2>&1; echo "pf tests: exit $? : $(tail -1 $O/t_pf.log)" > $O/summary.txt
for q in 2 3; do
echo "== throughput mode, at most $q QPs per SM" >> $O/summary.txt
for cfg in "128 100 100 0" "1024 100 100 0" "8192 100 100 0" "1024 50 50 10"; do
  QPB_KT_TWO=1 QPB200_MAXQPS=$q timeout 120 python scripts/kernel_times.py $cfg >> $O/summary.txt 2>&1
done
done
echo "== C4: 512 vs 256 threads" >> $O/summary.txt
timeout 120 python scripts/kernel_times.py 64 200 200 0 >> $O/summary.txt 2>&1
QPB200_NT512=0 timeout 120 python scripts/kernel_times.py 64 200 200 0 >> $O/summary.txt 2>&1
timeout 300 python scripts/c4_times.py >> $O/summary.txt 2>&1
QPB_BENCH_E2E=0 timeout 600 python bench.py --steps 24 --warmup 5 > $O/bench_v.json 2> $O/bench.err
QPB200_MAXQPS=2 QPB_BENCH_E2E=0 timeout 600 python bench.py --steps 24 --warmup 5 > $O/bench_v2.json 2>> $O/bench.err
QPB_BENCH_INFLIGHT=6 QPB_BENCH_E2E=0 timeout 600 python bench.py --steps 24 --warmup 5 > $O/bench_v_i6.json 2>> $O/bench.err
cat $O/summary.txt; tail -5 $O/t_pf.log; cat $O/bench_v.json $O/bench_v2.json $O/bench_v_i6.json; tail -3 $O/bench.err
