#!/bin/bash
TAG=${1:-r2s}
O=gpurun_out/$TAG
mkdir -p $O
for sp in 0 1; do for st in 24 48; do
QPB200_SETUP_PF=$sp QPB_BENCH_E2E=0 timeout 600 python bench.py --steps $st --warmup 5 > $O/bench_sp${sp}_k$st.json 2>> $O/bench.err
echo "SETUP_PF=$sp steps=$st: $(cat $O/bench_sp${sp}_k$st.json)"
done; done
timeout 600 python -m pytest tests/test_gpu_solution.py -m gpu -q -k lower_triangle 2>&1 | tail -2
