#!/bin/bash
mkdir -p gpurun_out/r2n
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k2_g2' -s 4 -c 1 -o gpurun_out/r2n/prof_k2 python bench.py --steps 2 --warmup 1 --blocks 2 --no-cpu > gpurun_out/r2n/ncu_k2.log 2>&1
tail -2 gpurun_out/r2n/ncu_k2.log
