#!/bin/bash
TAG=${1:-r2g}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "product_form_kernels" > $O/t_pf.log 2>&1; echo "pf tests: exit $? : $(tail -1 $O/t_pf.log)" > $O/summary.txt
for mode in 0 1 2; do
echo "== QPB200_PF=$mode" >> $O/summary.txt
for cfg in "128 100 100 0" "1024 100 100 0" "8192 100 100 0" "1024 50 50 10"; do
  QPB200_PF=$mode timeout 120 python scripts/kernel_times.py $cfg >> $O/summary.txt 2>&1
done
done
cat $O/summary.txt
