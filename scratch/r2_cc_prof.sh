#!/bin/bash
mkdir -p gpurun_out/r2m
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv64k' -s 4 -c 1 -o gpurun_out/r2m/prof_cc python bench.py --workload cfg3 --steps 1 --warmup 1 --blocks 2 --no-cpu > gpurun_out/r2m/ncu_cc.log 2>&1
tail -3 gpurun_out/r2m/ncu_cc.log
