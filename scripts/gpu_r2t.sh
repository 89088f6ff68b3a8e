#!/bin/bash
TAG=${1:-r2t}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kkt.py tests/test_gpu_layers.py -m gpu -q -x > $O/t.log 2>&1; echo "tests: exit $? : $(tail -1 $O/t.log)" > $O/summary.txt
for cfg in "128 100 100 0" "1024 100 100 0" "8192 100 100 0" "1024 50 50 10"; do QPB_KT_TWO=1 timeout 120 python scripts/kernel_times.py $cfg >> $O/summary.txt 2>&1; done
timeout 120 python scripts/kernel_times.py 64 200 200 0 >> $O/summary.txt 2>&1
QPB200_MAXQPS=2 QPB_KT_TWO=1 timeout 120 python scripts/kernel_times.py 8192 100 100 0 >> $O/summary.txt 2>&1
QPB_BENCH_E2E=0 timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_v.json 2> $O/bench.err
cat $O/summary.txt; cat $O/bench_v.json; tail -2 $O/bench.err
