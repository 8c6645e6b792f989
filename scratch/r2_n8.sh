#!/bin/bash
# round 2: 8-GPU lines of the sharded configurations + the host-to-device ceiling of the box
mkdir -p gpurun_out/r2h
N=$(nvidia-smi -L | wc -l)
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 "$@"; }
for w in cfg3 cfg5 cfg1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --workload $w --steps 10 --warmup 3 > gpurun_out/r2h/bench_${w}_n$N.json 2> gpurun_out/r2h/bench_${w}_n$N.err
  echo "$w rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2h/bench_${w}_n$N.json"))
    print("$w N=$N", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac %.3f" % d["roofline"]["frac"], d["config"]["combine"], str(d["hits_after_combine"])[:60])
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/r2h/bench_${w}_n$N.err").read()[-2500:])
PY
done
H2D_TOPO=1 H2D_MODE=plain run scratch/h2d_multi.py 2>/dev/null | tee gpurun_out/r2h/h2d_n$N.txt
for m in bind wc; do H2D_MODE=$m run scratch/h2d_multi.py 2>/dev/null | tee -a gpurun_out/r2h/h2d_n$N.txt; done
