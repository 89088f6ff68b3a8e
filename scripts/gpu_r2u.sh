#!/bin/bash
TAG=${1:-r2u}
O=gpurun_out/$TAG
mkdir -p $O
QPB200_LIB=$PWD/build/variants/lib_pairs.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or sweep" > $O/t.log 2>&1; echo "pairs variant tests: exit $? : $(tail -1 $O/t.log)" > $O/summary.txt
for lib in "" "$PWD/build/variants/lib_pairs.so"; do
echo "== lib=${lib:-product}" >> $O/summary.txt
for cfg in "128 100 100 0" "64 200 200 0"; do QPB200_LIB=$lib timeout 120 python scripts/kernel_times.py $cfg >> $O/summary.txt 2>&1; done
for cfg in "8192 100 100 0" "1024 50 50 10"; do QPB200_LIB=$lib QPB_KT_TWO=1 timeout 120 python scripts/kernel_times.py $cfg >> $O/summary.txt 2>&1; done
done
cat $O/summary.txt
