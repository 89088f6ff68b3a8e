#!/bin/bash
TAG=${1:-r2zb}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu: exit $? : $(tail -1 $O/pytest_gpu.log)" > $O/summary.txt
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
echo "== kernel times: throughput mode" > $O/kernel_times.txt
for cfg in "1024 100 100 0" "8192 100 100 0" "1024 50 50 10"; do QPB_KT_TWO=1 timeout 120 python scripts/kernel_times.py $cfg >> $O/kernel_times.txt 2>&1; done
timeout 120 python scripts/kernel_times.py 64 200 200 0 >> $O/kernel_times.txt 2>&1
cat $O/summary.txt $O/kernel_times.txt
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value %.0f ms/step %.3f e2e %.0f serial %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["detail"]["serial_value"]))
r=json.load(open("$O/bench_ref.json")); print("reference arm", r["value"])
PY
