#!/bin/bash
# round 2: GPU tests, then one bench line per BASELINE configuration (N = 1)
mkdir -p gpurun_out/r2d
python -m pytest tests -m gpu -x -q > gpurun_out/r2d/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d/pytest.log
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r2d/bench_$w.json 2> gpurun_out/r2d/bench_$w.err
  echo "$w rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2d/bench_$w.json"))
    print("$w", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac %.3f" % d["roofline"]["frac"], "blocks", d["config"]["blocks_per_step"], {k: round(v["ms_per_block"],3) for k,v in d["kernels"].items()})
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/r2d/bench_$w.err").read()[-1500:])
PY
done
