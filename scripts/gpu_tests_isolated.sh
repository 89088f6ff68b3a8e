#!/bin/bash
# GPU tests in separate processes (one CUDA fault must not poison the rest), then the whole suite in one process the
# way the driver runs it. Usage: bash scripts/gpu_tests_isolated.sh <outdir>
O=${1:-gpurun_out/tests}
mkdir -p $O
rm -f gpurun_out/parity_report.jsonl
run() { # name, pytest args...
  local name=$1; shift
  timeout 900 python -m pytest "$@" -q > $O/t_$name.log 2>&1
  echo "$name: exit $? : $(tail -1 $O/t_$name.log)" >> $O/summary.txt
}
: > $O/summary.txt
run golden_main tests/test_gpu_parity.py -m gpu -k "test_matches_reference_golden and not band"
for c in band_smem band_smem_eq band_setup band_setup_eq; do run $c tests/test_gpu_parity.py -m gpu -k "test_matches_reference_golden and $c]"; done
run sweep tests/test_gpu_parity.py -m gpu -k "sweep"
run rest tests/test_gpu_parity.py -m gpu -k "not test_matches_reference_golden and not sweep"
run solution tests/test_gpu_solution.py -m gpu
cp gpurun_out/parity_report.jsonl $O/ 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -x -q > $O/t_all_one_process.log 2>&1; echo "all-in-one: exit $? : $(tail -1 $O/t_all_one_process.log)" >> $O/summary.txt
cat $O/summary.txt
