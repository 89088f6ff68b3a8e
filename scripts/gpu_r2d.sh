#!/bin/bash
# phase accounting of the product-form kernel vs the shipped one (timing builds), racecheck, kernel times
TAG=${1:-r2d}
O=gpurun_out/$TAG
mkdir -p $O
QPB200_TIMING_LIB=$PWD/build/timing/t_r2.so QPB200_PF=1 timeout 120 python scripts/phase_timing.py > $O/phase_pf.log 2>&1
QPB200_TIMING_LIB=$PWD/build/timing/t_r2.so timeout 120 python scripts/phase_timing.py > $O/phase_std.log 2>&1
QPB200_TIMING_LIB=$PWD/build/timing/t_r2.so QPB200_PF=1 timeout 120 python scripts/phase_timing.py 64 200 200 0 > $O/phase_pf_c4.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "product_form_kernels and c3_b64]" > $O/racecheck_c3.log 2>&1
echo "racecheck c3_b64: $(grep 'RACECHECK SUMMARY' $O/racecheck_c3.log | tail -1)" > $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "product_form" > $O/t_pf.log 2>&1; echo "pf tests: exit $? : $(tail -1 $O/t_pf.log)" >> $O/summary.txt
for cfg in "128 100 100 0" "1024 100 100 0" "64 200 200 0"; do
  QPB200_PF=1 timeout 120 python scripts/kernel_times.py $cfg >> $O/summary.txt 2>&1
done
cat $O/summary.txt $O/phase_pf.log; tail -45 $O/phase_std.log; cat $O/phase_pf_c4.log
