#!/bin/bash
TAG=${1:-r2r}
O=gpurun_out/$TAG
mkdir -p $O
QPB200_SETUP_PF=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_matches_reference_golden or test_pre_factor" > $O/t.log 2>&1; echo "golden with setup_pf everywhere: exit $? : $(tail -1 $O/t.log)" > $O/summary.txt
for sp in 1 0; do
echo "== QPB200_SETUP_PF=$sp (throughput mode)" >> $O/summary.txt
for cfg in "128 100 100 0" "8192 100 100 0" "1024 50 50 10" "64 200 200 0"; do
  QPB200_SETUP_PF=$sp QPB_KT_TWO=1 timeout 120 python scripts/kernel_times.py $cfg >> $O/summary.txt 2>&1
done
done
cat $O/summary.txt
