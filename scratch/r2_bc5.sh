#!/bin/bash
mkdir -p gpurun_out/r2s
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "fold or pipeline or golden or stream or subint" 2>&1 | tail -2
for c in cfg5 cfg4; do
python bench.py --workload $c --steps 10 --warmup 3 --no-cpu > gpurun_out/r2s/bench_${c}_fin.json 2> gpurun_out/r2s/bench_${c}_fin.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2s/bench_${c}_fin.json"))
print(round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v["ms_per_block"], 4) for k, v in d["kernels"].items()})
PY
done
