#!/bin/bash
TAG=${1:-r2l}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
QPTH_B200_MODE=throughput timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(setup|forward|kkt)' -s 3 -c 3 \
    -o $O/hot_kernels_b2048 -f python scripts/prof_one.py 2048 100 100 0 > $O/ncu_b2048.log 2>&1
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value %.0f ms/step %.3f e2e %.0f (%s) serial %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["launch"][:20], d["detail"]["serial_value"]))
print("e2e windows", d["e2e"]["windows_ms"], "default opts", d["e2e"]["default_options"]["value"])
print("c4", d["detail"].get("c4"))
PY
tail -3 $O/bench.err
