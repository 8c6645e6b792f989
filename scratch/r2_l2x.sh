#!/bin/bash
mkdir -p gpurun_out/r2m
L=$PWD/dspsr_b200
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg3 or cluster" 2>&1 | tail -2
B200_LIB=$L/libb200dsp_dsm.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg3 or cluster" 2>&1 | tail -2
for tag in sw2 new dsm new2; do
  lib=$L/libb200dsp_${tag:0:3}.so; [ ${tag:0:3} = new ] && lib=$L/libb200dsp.so
  B200_LIB=$lib timeout 300 python bench.py --workload cfg3 --steps 4 --warmup 2 --no-cpu > gpurun_out/r2m/abl_$tag.json 2> gpurun_out/r2m/abl_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2m/abl_$tag.json"))
    print("$tag", round(d["value"]), {k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2m/abl_$tag.err").read()[-600:])
PY
done
