#!/bin/bash
# final-state multi-GPU check on 2 GPUs: NCCL parity tests of all sharding modes + native host, one more parity test in
# the same session (regression check of the earlier 'invalid device ordinal'), cfg1 / cfg3 bench lines at N = 2
out=gpurun_out/r2q; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -rs -k "sharded or native_multi or execute_obs" > $out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $out/pytest_multi.log
tail -12 $out/pytest_multi.log
N=2
for w in cfg1 cfg3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --workload $w --steps 10 --warmup 3 --no-cpu > $out/bench_${w}_n$N.json 2> $out/bench_${w}_n$N.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_${w}_n$N.json"))
    print("$w N=$N", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac %.3f" % d["roofline"]["frac"], d["config"]["combine"], d["hits_after_combine"])
except Exception as e:
    print("$w failed", e); print(open("$out/bench_${w}_n$N.err").read()[-2500:])
PY
done
