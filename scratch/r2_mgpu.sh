#!/bin/bash
# round 2: multi-GPU parity (NCCL) + a 2-GPU bench line per sharding mode
mkdir -p gpurun_out/r2e
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs > gpurun_out/r2e/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2e/pytest_multi.log
N=$(nvidia-smi -L | wc -l)
for w in cfg1 cfg3 cfg5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --workload $w --steps 10 --warmup 3 > gpurun_out/r2e/bench_${w}_n$N.json 2> gpurun_out/r2e/bench_${w}_n$N.err
  echo "$w rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2e/bench_${w}_n$N.json"))
    print("$w N=$N", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac %.3f" % d["roofline"]["frac"], d["config"]["combine"], d["hits_after_combine"])
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/r2e/bench_${w}_n$N.err").read()[-2500:])
PY
done
./scratch/phasebench/phasebench 64 > gpurun_out/r2e/phasebench.txt 2>&1; cat gpurun_out/r2e/phasebench.txt
