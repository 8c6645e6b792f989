#!/bin/bash
mkdir -p gpurun_out/r2c
for pp in 0 1 2; do
  B200_K2_PP=$pp timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "cfg1 or pipeline_small" 2>&1 | tail -2 > gpurun_out/r2c/pytest_pp$pp.log
  B200_K2_PP=$pp python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/r2c/bench_pp$pp.json 2> gpurun_out/r2c/bench_pp$pp.err
done
for pp in 0 1 2; do tail -1 gpurun_out/r2c/pytest_pp$pp.log; python -c "
import json; d=json.load(open('gpurun_out/r2c/bench_pp$pp.json')); print('pp$pp', round(d['value']), {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"; done
