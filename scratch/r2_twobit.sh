#!/bin/bash
# two-bit unpack fused into the generic K1: parity of every two-bit test, then cfg2 with and without (developer build)
mkdir -p gpurun_out/r2r
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "twobit or two_bit or cfg2 or unpack" 2>&1 | tail -2
L=$PWD/dspsr_b200/libb200dsp_dev.so
for spec in unfused:0 fused:1 unfused2:0 fused2:1; do
  tag=${spec%%:*}; v=${spec#*:}
  B200_LIB=$L B200_TWOBIT_FUSED=$v python bench.py --workload cfg2 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2r/bench_$tag.json 2> gpurun_out/r2r/bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2r/bench_$tag.json"))
    print("$tag", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms_per_block"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2r/bench_$tag.err").read()[-800:])
PY
done
