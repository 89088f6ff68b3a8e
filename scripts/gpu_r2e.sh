#!/bin/bash
TAG=${1:-r2e}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "product_form" > $O/t_pf.log 2>&1; echo "pf tests: exit $? : $(tail -1 $O/t_pf.log)" > $O/summary.txt
for cfg in "128 100 100 0" "1024 100 100 0" "1024 50 50 10" "64 200 200 0"; do
  QPB200_PF=1 timeout 120 python scripts/kernel_times.py $cfg >> $O/summary.txt 2>&1
done
QPB200_TIMING_LIB=$PWD/build/timing/t_r2.so QPB200_PF=1 timeout 120 python scripts/phase_timing.py > $O/phase_pf.log 2>&1
QPB200_TIMING_LIB=$PWD/build/timing/t_r2.so QPB200_PF=1 timeout 120 python scripts/phase_timing.py 64 200 200 0 > $O/phase_pf_c4.log 2>&1
cat $O/summary.txt $O/phase_pf.log; tail -32 $O/phase_pf_c4.log
