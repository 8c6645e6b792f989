#!/bin/bash
# final-state 8-GPU lines of cfg3 (one-kernel convolution path, channel shards) and cfg1 (time shards, NCCL reduce)
mkdir -p gpurun_out/r2s
N=$(nvidia-smi -L | wc -l)
for w in cfg3 cfg1; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r2s/bench_${w}_n$N.json 2> gpurun_out/r2s/bench_${w}_n$N.err
  echo "$w rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2s/bench_${w}_n$N.json"))
    print("$w N=$N", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac %.3f" % d["roofline"]["frac"], d["config"]["combine"], str(d["hits_after_combine"])[:60])
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/r2s/bench_${w}_n$N.err").read()[-2500:])
PY
done
