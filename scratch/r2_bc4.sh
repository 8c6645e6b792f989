#!/bin/bash
# K1c merged 32-bit loads + L2 prefetch of the following tile (column kernels): parity, then cfg4 A/B over the distance
mkdir -p gpurun_out/r2s
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_conv or cfg4" 2>&1 | tail -2
L=$PWD/dspsr_b200/libb200dsp_dev.so
run() {  # tag workload env...
  tag=$1; wl=$2; shift 2
  env B200_LIB=$L "$@" python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu > gpurun_out/r2s/bench_$tag.json 2> gpurun_out/r2s/bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2s/bench_$tag.json"))
    print("$tag", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms_per_block"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2s/bench_$tag.err").read()[-800:])
PY
}
run cfg4_ahead0 cfg4 B200_BC_AHEAD=0
run cfg4_ahead1 cfg4 B200_BC_AHEAD=1
run cfg4_ahead2 cfg4 B200_BC_AHEAD=2
