#!/bin/bash
# phase ablations of the cluster convolution kernel (results are wrong by construction: timing only)
mkdir -p gpurun_out/r2m
L=$PWD/dspsr_b200
for tag in full cc1 cc2 cc4 cc7; do
  lib=$L/libb200dsp_$tag.so; [ $tag = full ] && lib=$L/libb200dsp.so
  B200_LIB=$lib timeout 300 python bench.py --workload cfg3 --steps 4 --warmup 2 --no-cpu > gpurun_out/r2m/abl_$tag.json 2> gpurun_out/r2m/abl_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2m/abl_$tag.json"))
    print("$tag", round(d["value"]), {k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2m/abl_$tag.err").read()[-800:])
PY
done
