#!/bin/bash
# Round-2 first GPU pass: parity (all gpu tests, no -x), A/B of the kernel families (coop x trsv16), in-flight sweep
# of the bench, both bench arms, phase accounting, ncu launch list + full capture.
TAG=${1:-r2a}
O=gpurun_out/$TAG
mkdir -p $O
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt; lscpu | head -20 >> $O/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
cp gpurun_out/parity_report.jsonl $O/ 2>/dev/null
echo "== kernel times" > $O/kernel_times.log
for coop in 0 1; do
  for lib in "" build/variants/lib_trsv8.so; do
    echo "-- coop=$coop lib=${lib:-product(trsv16)}" >> $O/kernel_times.log
    for cfg in "128 100 100 0" "1024 100 100 0" "8192 100 100 0" "1024 50 50 10"; do
      QPB200_COOP=$coop QPB200_LIB=${lib:+$PWD/$lib} timeout 120 python scripts/kernel_times.py $cfg >> $O/kernel_times.log 2>&1
    done
  done
done
timeout 120 python scripts/kernel_times.py 64 200 200 0 >> $O/kernel_times.log 2>&1
timeout 120 python scripts/kernel_times.py 4096 10 5 0 >> $O/kernel_times.log 2>&1
for coop in 0 1; do
  QPB200_COOP=$coop QPB200_TIMING_LIB=$PWD/build/timing/t_r2.so timeout 120 python scripts/phase_timing.py > $O/phase_coop$coop.log 2>&1
done
# in-flight sweep (resident value only)
for coop in 0 1; do for inf in 3 4 6; do
  QPB200_COOP=$coop QPB_BENCH_INFLIGHT=$inf QPB_BENCH_E2E=0 QPB_BENCH_CPU=0 QPB_BENCH_E2E_DEFAULT=0 QPB_BENCH_C4=0 QPB_BENCH_REFCUDA=0 timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_c${coop}_i$inf.json 2> $O/bench_c${coop}_i$inf.err
done; done
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench*.json")):
    try:
        d=json.load(open(f)); dt=d.get("detail",{})
        print(f, "value %.0f (%.3f ms/step, serial %.3f) e2e %.0f" % (d["value"], d["ms_per_step"], dt.get("serial_ms_per_step",0), d["e2e"]["value"]))
    except Exception as e: print(f, "unreadable", e)
PY
if [ -z "$SKIP_NCU" ]; then
QPB_BENCH_CPU=0 QPB_BENCH_E2E_DEFAULT=0 QPB_BENCH_C4=0 QPB_BENCH_REFCUDA=0 QPB_BENCH_MAX_SETTLE=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(setup|forward|kkt)' -s 3 -c 3 \
    -o $O/hot_kernels -f python scripts/prof_one.py > $O/ncu_full.log 2>&1
fi
ls -la $O
tail -5 $O/pytest_gpu.log; cat $O/kernel_times.log
