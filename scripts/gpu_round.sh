#!/bin/bash
# One GPU-box pass: parity tests, both bench arms, kernel-only times, phase timing, ncu launch list + full capture.
# Everything lands in gpurun_out/ (merged back by gpurun). Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag]'
TAG=${1:-r1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
fi
timeout 120 python scripts/kernel_times.py 128 100 100 0 > $O/kernel_times.log 2>&1
timeout 120 python scripts/kernel_times.py 1024 50 50 10 >> $O/kernel_times.log 2>&1
timeout 120 python scripts/kernel_times.py 1024 100 100 0 >> $O/kernel_times.log 2>&1
timeout 120 python scripts/phase_timing.py > $O/phase_timing.log 2>&1
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
if [ -z "$SKIP_REF" ]; then
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
fi
if [ -z "$SKIP_NCU" ]; then
# launch list of the bench command (cold-cache, serialised: shares only)
QPB_BENCH_MAX_SETTLE=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 > $O/bench_under_ncu.log 2>&1
# full capture of the three hot kernels (second repetition = warm)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(setup|forward|kkt)' -s 3 -c 3 \
    -o $O/hot_kernels -f python scripts/prof_one.py > $O/ncu_full.log 2>&1
fi
ls -la $O
tail -3 $O/pytest_gpu.log; cat $O/kernel_times.log; head -c 3000 $O/bench.json
