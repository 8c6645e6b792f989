#!/bin/bash
mkdir -p gpurun_out/r2r
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "filterbank or cfg2 or twobit or stream or short_channel or golden" 2>&1 | tail -2
L=$PWD/dspsr_b200/libb200dsp_dev.so
for spec in old:0 new:1 old2:0 new2:1; do
  tag=${spec%%:*}; v=${spec#*:}
  B200_LIB=$L B200_K3_SMALL=$v python bench.py --workload cfg2 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2r/bench_k3s_$tag.json 2> gpurun_out/r2r/bench_k3s_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2r/bench_k3s_$tag.json"))
    print("$tag", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms_per_block"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2r/bench_k3s_$tag.err").read()[-800:])
PY
done
