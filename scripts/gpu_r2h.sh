#!/bin/bash
# full GPU suite + both bench arms with the product-form kernels as the default
TAG=${1:-r2h}
O=gpurun_out/$TAG
mkdir -p $O
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu: exit $? : $(tail -1 $O/pytest_gpu.log)" > $O/summary.txt
cp gpurun_out/parity_report.jsonl $O/ 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
QPB_BENCH_INFLIGHT=6 QPB_BENCH_E2E=0 timeout 600 python bench.py --steps 24 --warmup 5 > $O/bench_i6.json 2>> $O/bench.err
QPB_BENCH_INFLIGHT=3 QPB_BENCH_E2E=0 timeout 600 python bench.py --steps 24 --warmup 5 > $O/bench_i3.json 2>> $O/bench.err
QPB_BENCH_MODE=latency QPB_BENCH_E2E=0 timeout 600 python bench.py --steps 24 --warmup 5 > $O/bench_lat.json 2>> $O/bench.err
python __graft_entry__.py --smoke > $O/smoke.log 2>&1
cat $O/summary.txt; tail -5 $O/pytest_gpu.log; cat $O/smoke.log | tail -2
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value %.0f ms/step %.3f e2e %.0f serial %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["detail"]["serial_value"]))
print("e2e windows", d["e2e"]["windows_ms"], "default opts", d["e2e"]["default_options"]["value"])
print("kernel alone", d["detail"]["kernel_ms_alone"], d["detail"]["solve_kernels"])
print("c4", d["detail"].get("c4"))
print("cpu", d["cpu_baseline"]["value"], "refcuda", d.get("reference_cuda",{}).get("value"))
for f in ("bench_i6","bench_i3","bench_lat"):
    try:
        x=json.load(open("$O/%s.json"%f)); print(f, "value %.0f serial %.3f ms" % (x["value"], x["detail"]["serial_ms_per_step"]))
    except Exception as e: print(f, "unreadable", e)
PY
tail -3 $O/bench.err
