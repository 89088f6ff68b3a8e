O=gpurun_out/r2zf; mkdir -p $O
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest -m gpu: exit $? : $(tail -1 $O/pytest_gpu.log)" > $O/summary.txt
cp gpurun_out/parity_report.jsonl $O/ 2>/dev/null
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
python __graft_entry__.py --smoke > $O/smoke.log 2>&1
cat $O/summary.txt; tail -1 $O/smoke.log
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value %.0f ms/step %.3f e2e %.0f serial %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["detail"]["serial_value"]))
print("default", d["e2e"]["default_options"]["value"], "lazy", d["e2e"]["default_options"].get("lazy_checks"))
print("c4", d["detail"].get("c4"))
r=json.load(open("$O/bench_ref.json")); print("reference arm", r["value"])
PY
tail -3 $O/bench.err
