#!/bin/bash
# final round-2 state: parity suite, bench lines of the five configurations, ncu launch list, ncu --set full captures
# (cfg1 kernels, cfg3 one-kernel path, cfg4 long-transform kernels).  usage: scratch/r2_profile2.sh TAG
tag=${1:-r2q}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
python bench.py --steps 20 --warmup 3 > $out/bench_cfg1.json 2> $out/bench_cfg1.err
for c in cfg2 cfg3 cfg4 cfg5; do
  python bench.py --workload $c --steps 10 --warmup 3 --no-cpu > $out/bench_$c.json 2> $out/bench_$c.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --blocks 2 --no-cpu > $out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k1_c2|k2_g2|k2_r32|k3_c2' -s 12 -c 3 -o $out/prof_full python bench.py --steps 2 --warmup 1 --blocks 2 --no-cpu > $out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_bc_cols|k_bc_rows|k_expand' -s 8 -c 4 -o $out/prof_long python bench.py --workload cfg4 --steps 1 --warmup 1 --blocks 2 --no-cpu > $out/ncu_long.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_cfg4.csv python bench.py --workload cfg4 --steps 2 --warmup 1 --blocks 2 --no-cpu > $out/bench_cfg4_under_ncu.log 2>&1
tail -3 $out/pytest.log; for c in cfg1 cfg2 cfg3 cfg4 cfg5; do cut -c1-160 $out/bench_$c.json; echo; done
