#!/bin/bash
# Round-2 pass C: product-form / staircase kernels (qp_pf.cuh): parity, memcheck, kernel times against the shipped kernels.
TAG=${1:-r2c}
O=gpurun_out/$TAG
mkdir -p $O
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "product_form_kernels" > $O/t_pf_golden.log 2>&1; echo "pf_golden: exit $? : $(tail -1 $O/t_pf_golden.log)" > $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sweep_product_form" > $O/t_pf_sweep.log 2>&1; echo "pf_sweep: exit $? : $(tail -1 $O/t_pf_sweep.log)" >> $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "test_matches_reference_golden" > $O/t_golden.log 2>&1; echo "golden(default plans): exit $? : $(tail -1 $O/t_golden.log)" >> $O/summary.txt
for c in c3_b64 c4; do
QPB200_PF=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 6 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "product_form_kernels and $c]" > $O/memcheck_$c.log 2>&1
echo "memcheck $c: $(grep 'ERROR SUMMARY' $O/memcheck_$c.log | tail -1)" >> $O/summary.txt
done
timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "product_form_kernels and c3_b64]" > $O/racecheck_c3.log 2>&1
echo "racecheck c3_b64: $(grep 'RACECHECK SUMMARY' $O/racecheck_c3.log | tail -1)" >> $O/summary.txt
echo "== kernel times, shipped kernels" > $O/kernel_times.log
for cfg in "128 100 100 0" "1024 100 100 0" "1024 50 50 10" "64 200 200 0"; do
  timeout 120 python scripts/kernel_times.py $cfg >> $O/kernel_times.log 2>&1
done
echo "== kernel times, QPB200_PF=1" >> $O/kernel_times.log
for cfg in "128 100 100 0" "1024 100 100 0" "1024 50 50 10" "64 200 200 0" "1024 20 120 0"; do
  QPB200_PF=1 timeout 120 python scripts/kernel_times.py $cfg >> $O/kernel_times.log 2>&1
done
timeout 300 python scripts/c4_times.py > $O/c4_times.log 2>&1
cp gpurun_out/parity_report.jsonl $O/ 2>/dev/null
cat $O/summary.txt $O/kernel_times.log $O/c4_times.log
grep -B2 -A12 "Invalid\|Error\|error" $O/memcheck_c3_b64.log | head -60
tail -30 $O/t_pf_golden.log
