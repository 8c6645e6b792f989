#!/bin/bash
# round-2 first look: parity suite, baseline number, small-batch (L2-resident ring) sweep of the round-1 kernels
mkdir -p gpurun_out/r2a
nvidia-smi -L > gpurun_out/r2a/gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a/bench_base.json 2> gpurun_out/r2a/bench_base.err
for b in 1 2 3 4; do
  python bench.py --steps 6 --warmup 3 --no-cpu --parts 24 --batch $b > gpurun_out/r2a/bench_b$b.json 2> gpurun_out/r2a/bench_b$b.err
done
for b in 1 2; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'k1_c2|k2_r32|k3_c2' -s 30 -c 12 --csv --log-file gpurun_out/r2a/ncu_b$b.csv python bench.py --steps 2 --warmup 3 --no-cpu --parts 8 --batch $b > /dev/null 2>&1
done
tail -3 gpurun_out/r2a/pytest.log
