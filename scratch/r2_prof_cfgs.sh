#!/bin/bash
# ncu --set full of the three main kernels of cfg2 / cfg3 / cfg4 (generic kernels), one capture each
out=gpurun_out/r2l; mkdir -p $out
for c in cfg3 cfg4 cfg2; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_cols|k_rows|k_chan|k1_|k2_|k3_' -s 9 -c 3 -o $out/prof_$c python bench.py --workload $c --steps 1 --warmup 1 --blocks 2 --no-cpu > $out/ncu_$c.log 2>&1
done
ls -la $out
