#!/bin/bash
# Round-2 final evidence pass: full suite, both bench arms, ncu launch list of the bench command, ncu --set full of the hot
# kernels (latency variant at B=128, three-per-SM variant at B=2048, C4), phase accounting, kernel times.
TAG=${1:-r2z}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu: exit $? : $(tail -1 $O/pytest_gpu.log)" > $O/summary.txt
cp gpurun_out/parity_report.jsonl $O/ 2>/dev/null
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
echo "== kernel times: latency mode" > $O/kernel_times.txt
for cfg in "128 100 100 0" "1024 50 50 10" "64 200 200 0" "4096 10 5 0"; do timeout 120 python scripts/kernel_times.py $cfg >> $O/kernel_times.txt 2>&1; done
echo "== kernel times: throughput mode (three QPs per SM where the shape allows)" >> $O/kernel_times.txt
for cfg in "128 100 100 0" "1024 100 100 0" "8192 100 100 0" "1024 50 50 10"; do QPB_KT_TWO=1 timeout 120 python scripts/kernel_times.py $cfg >> $O/kernel_times.txt 2>&1; done
timeout 300 python scripts/c4_times.py >> $O/kernel_times.txt 2>&1
QPB200_TIMING_LIB=$PWD/build/timing/t_r2.so timeout 120 python scripts/phase_timing.py > $O/phase_latency.txt 2>&1
QPB_BENCH_MAX_SETTLE=8 QPB_BENCH_CPU=0 QPB_BENCH_C4=0 QPB_BENCH_REFCUDA=0 QPB_BENCH_E2E_DEFAULT=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv \
    --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 > $O/bench_under_ncu.log 2>&1
QPTH_B200_MODE=latency timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(setup|forward|kkt)' -s 3 -c 3 \
    -o $O/hot_latency_b128 -f python scripts/prof_one.py > $O/ncu1.log 2>&1
QPTH_B200_MODE=throughput timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(setup|forward|kkt)' -s 3 -c 3 \
    -o $O/hot_throughput_b2048 -f python scripts/prof_one.py 2048 100 100 0 > $O/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_(setup|forward|kkt)' -s 3 -c 3 -o $O/hot_c4 -f python scripts/prof_one.py 64 200 200 0 > $O/ncu3.log 2>&1
python __graft_entry__.py --smoke > $O/smoke.log 2>&1
cat $O/summary.txt; tail -3 $O/pytest_gpu.log; cat $O/kernel_times.txt; tail -2 $O/smoke.log
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value %.0f ms/step %.3f e2e %.0f serial %.0f launch %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["detail"]["serial_value"], d["detail"]["launch"]))
print("e2e windows", d["e2e"]["windows_ms"], "default", d["e2e"]["default_options"]["value"], d["e2e"]["launch"][:30])
print("c4", d["detail"].get("c4")); print("cpu", d["cpu_baseline"]["value"], "refcuda", d.get("reference_cuda",{}).get("value"))
r=json.load(open("$O/bench_ref.json")); print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
tail -3 $O/bench.err
