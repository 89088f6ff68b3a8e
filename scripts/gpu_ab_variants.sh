#!/bin/bash
# A/B of kernel build variants on ONE box, two repetitions each (the method behind profiles/r2y_chunk_sweepbar_ab.txt):
# build them with scripts/build_variants.sh real "NAME:-DFLAG=..." ..., list the NAMEs in the loop below.
TAG=${1:-ab}
O=gpurun_out/$TAG
mkdir -p $O
: > $O/summary.txt
for rep in 1 2; do
for v in c0s0 c0s1 c1s0 c2s0 c2s1; do
echo "== $v (rep $rep)" >> $O/summary.txt
QPB200_LIB=$PWD/build/variants/lib_$v.so timeout 120 python scripts/kernel_times.py 128 100 100 0 >> $O/summary.txt 2>&1
QPB200_LIB=$PWD/build/variants/lib_$v.so QPB_KT_TWO=1 timeout 120 python scripts/kernel_times.py 128 100 100 0 >> $O/summary.txt 2>&1
QPB200_LIB=$PWD/build/variants/lib_$v.so QPB_KT_TWO=1 timeout 120 python scripts/kernel_times.py 8192 100 100 0 >> $O/summary.txt 2>&1
QPB200_LIB=$PWD/build/variants/lib_$v.so timeout 120 python scripts/kernel_times.py 64 200 200 0 >> $O/summary.txt 2>&1
done; done
grep -E "==|forward" $O/summary.txt | sed -e 's/ fast=.: setup [0-9.]* us,//' -e 's/ -> .*//'
