#!/bin/bash
TAG=${1:-r2p}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_layers.py -m gpu -q > $O/t_layers.log 2>&1; echo "layers: exit $? : $(tail -1 $O/t_layers.log)" > $O/summary.txt
QPB_BENCH_E2E=0 timeout 600 python bench.py --steps 24 --warmup 5 > $O/bench_v.json 2> $O/bench.err
cat $O/summary.txt; tail -15 $O/t_layers.log; cat $O/bench_v.json; tail -5 $O/bench.err
