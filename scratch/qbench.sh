#!/bin/bash
# quick GPU check: cfg1-shaped parity tests + bench; usage: scratch/qbench.sh <tag> [bench args]
tag=$1; shift
python -m pytest tests -m gpu -x -q -k "cfg1 or golden or pipeline" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value %.0f e2e %.0f MS/s  path frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline_path"]["frac"]))
for k,v in d["kernels"].items(): print("  %-9s %.4f ms/launch  %5.0f GB/s" % (k, v["ms_per_launch"], v.get("gbs",0)))
PY
tail -3 gpurun_out/bench_$tag.err
