#!/bin/bash
# bin-plan kernel with chunked hit counting (parity: every fold test compares the hit counts exactly), and the
# factorisation of cfg4's 2^22 points: 2048 x 2048 against 1024 x 4096 (dev build, B200_LGP_CAP)
mkdir -p gpurun_out/r2s
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "fold or pipeline or golden or long_conv or cluster" 2>&1 | tail -3
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
run cfg4_p2048 cfg4 B200_LGP_CAP=11
run cfg4_p1024 cfg4 B200_LGP_CAP=10
run cfg1 cfg1 B200_LGP_CAP=11
