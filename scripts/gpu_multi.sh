#!/bin/bash
# multi-GPU pass (gpurun --gpus N): NCCL tests + both bench arms under torchrun with the config-5 job
N=${1:-2}; TAG=${2:-r2m$N}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parallel.py -m gpu -q > $O/t_parallel.log 2>&1; echo "nccl tests ($N GPUs): exit $? : $(tail -1 $O/t_parallel.log)" > $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "bench exit $?" >> $O/summary.txt
cat $O/summary.txt; tail -5 $O/t_parallel.log
python - <<PY
import json
d=json.load(open("$O/bench_n$N.json"))
print("N=%d value %.0f e2e %.0f" % (d["n_gpus"], d["value"], d["e2e"]["value"]))
print("c5", json.dumps(d["detail"].get("c5")))
PY
tail -3 $O/bench_n$N.err
