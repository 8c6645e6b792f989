#!/bin/bash
# back-to-back A/B on one box: scratch/ab.sh tag[:ENV=VAL[,ENV=VAL]] ...   (B200_LIB selects a library build)
for spec in "$@"; do
  tag=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$tag.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("$tag", round(d["value"]), {k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
PY
done
