#!/bin/bash
# round-2 profile of the current state: parity suite, cfg1 bench line, ncu launch list and one ncu --set full capture
# usage: scratch/r2_profile.sh TAG
tag=${1:-r2j}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
python bench.py --steps 20 --warmup 3 > $out/bench_cfg1.json 2> $out/bench_cfg1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --blocks 2 --no-cpu > $out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k1_c2|k2_g2|k2_r32|k3_c2' -s 12 -c 3 -o $out/prof_full python bench.py --steps 2 --warmup 1 --blocks 2 --no-cpu > $out/ncu_full.log 2>&1
tail -3 $out/pytest.log; cut -c1-300 $out/bench_cfg1.json
