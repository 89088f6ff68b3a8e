#!/bin/bash
# A/B of kernel build variants on ONE box, two repetitions each (the method behind profiles/r2y_chunk_sweepbar_ab.txt):
# build them with scripts/build_variants.sh real "NAME:-DFLAG=..." ..., then  bash scripts/gpu_ab_variants.sh <tag> NAME ...
TAG=${1:-ab}; shift
O=gpurun_out/$TAG
mkdir -p $O
: > $O/summary.txt
for v in "$@"; do
QPB200_LIB=$PWD/build/variants/lib_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or families" > $O/t_$v.log 2>&1; echo "tests $v: exit $? : $(tail -1 $O/t_$v.log)" >> $O/summary.txt
done
for rep in 1 2; do
for v in "$@"; do
echo "== $v (rep $rep)" >> $O/summary.txt
QPB200_LIB=$PWD/build/variants/lib_$v.so timeout 120 python scripts/kernel_times.py 128 100 100 0 >> $O/summary.txt 2>&1
QPB200_LIB=$PWD/build/variants/lib_$v.so QPB_KT_TWO=1 timeout 120 python scripts/kernel_times.py 128 100 100 0 >> $O/summary.txt 2>&1
QPB200_LIB=$PWD/build/variants/lib_$v.so QPB_KT_TWO=1 timeout 120 python scripts/kernel_times.py 8192 100 100 0 >> $O/summary.txt 2>&1
QPB200_LIB=$PWD/build/variants/lib_$v.so timeout 120 python scripts/kernel_times.py 64 200 200 0 >> $O/summary.txt 2>&1
done; done
grep -E "tests|==|forward" $O/summary.txt | sed -e 's/ fast=.: setup [0-9.]* us,//' -e 's/ -> .*//'
