#!/bin/bash
mkdir -p gpurun_out/r2m
L=$PWD/dspsr_b200
B200_LIB=$L/libb200dsp_cc8.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg3 or 65536 or convolution" 2>&1 | tail -2
for tag in full cc8; do
  lib=$L/libb200dsp_$tag.so; [ $tag = full ] && lib=$L/libb200dsp.so
  B200_LIB=$lib timeout 300 python bench.py --workload cfg3 --steps 4 --warmup 2 --no-cpu > gpurun_out/r2m/abl_$tag.json 2> gpurun_out/r2m/abl_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2m/abl_$tag.json"))
print("$tag", round(d["value"]), {k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
PY
done
