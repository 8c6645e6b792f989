#!/bin/bash
mkdir -p gpurun_out/r2s
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_conv or cfg4" 2>&1 | tail -2
python bench.py --workload cfg4 --steps 6 --warmup 3 --no-cpu > gpurun_out/r2s/bench_cfg4_new.json 2> gpurun_out/r2s/bench_cfg4_new.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2s/bench_cfg4_new.json"))
print(round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms_per_block"], 4) for k, v in d["kernels"].items()})
PY
