#!/bin/bash
# ncu --set full of the three long-transform kernels and the bin-plan kernel at cfg4
out=gpurun_out/r2s; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bc_cols|k_bc_rows|k_expand' -s 8 -c 4 -o $out/prof_bc python bench.py --workload cfg4 --steps 1 --warmup 1 --blocks 2 --no-cpu > $out/ncu_bc.log 2>&1
ls -la $out/*.ncu-rep
