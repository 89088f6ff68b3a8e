#!/bin/bash
# Round-2 evidence pass: full suite, both bench arms, ncu launch list of the bench command, ncu --set full of the hot kernels
# (throughput variant = what the bench's timed region runs, and latency variant), phase accounting.
TAG=${1:-r2j}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu: exit $? : $(tail -1 $O/pytest_gpu.log)" > $O/summary.txt
cp gpurun_out/parity_report.jsonl $O/ 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
QPB_BENCH_DQ_BAND=0 QPB_BENCH_CPU=0 QPB_BENCH_C4=0 QPB_BENCH_REFCUDA=0 QPB_BENCH_E2E_DEFAULT=0 timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_fulldq.json 2>> $O/bench.err
QPB_BENCH_MAX_SETTLE=8 QPB_BENCH_CPU=0 QPB_BENCH_C4=0 QPB_BENCH_REFCUDA=0 QPB_BENCH_E2E_DEFAULT=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 > $O/bench_under_ncu.log 2>&1
QPTH_B200_MODE=throughput timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(setup|forward|kkt)' -s 3 -c 3 \
    -o $O/hot_kernels_two -f python scripts/prof_one.py > $O/ncu_full_two.log 2>&1
QPTH_B200_MODE=latency timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(forward)' -s 1 -c 1 \
    -o $O/hot_kernels_one -f python scripts/prof_one.py > $O/ncu_full_one.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_(setup|forward|kkt)' -s 3 -c 3 -o $O/hot_kernels_c4 -f python scripts/prof_one.py 64 200 200 0 > $O/ncu_full_c4.log 2>&1
cat $O/summary.txt; tail -3 $O/pytest_gpu.log
python - <<PY
import json
for f in ("bench","bench_fulldq"):
    d=json.load(open("$O/%s.json"%f))
    print(f, "value %.0f ms/step %.3f e2e %.0f serial %.0f d2h %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["detail"]["serial_value"], d["e2e"]["d2h_bytes_per_step"]))
    print("  e2e windows", d["e2e"]["windows_ms"])
d=json.load(open("$O/bench_ref.json")); print("reference arm", d["value"], d["cpu_baseline"]["cores"])
PY
ls -la $O
