#!/bin/bash
mkdir -p gpurun_out/r2m
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg3 or 65536 or convolution or cluster or one_kernel" 2>&1 | tail -1; done
python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2m/bench_cfg3_new.json 2> gpurun_out/r2m/bench_cfg3_new.err
python -c "
import json; d=json.load(open('gpurun_out/r2m/bench_cfg3_new.json')); print('cfg3', round(d['value']), {k: round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_conv64k -s 4 -c 2 --csv --log-file gpurun_out/r2m/ncu_cc_dram.csv python bench.py --workload cfg3 --steps 1 --warmup 1 --blocks 2 --no-cpu > /dev/null 2>&1
grep -E "k_conv64k" gpurun_out/r2m/ncu_cc_dram.csv | cut -d, -f5,13- | head -8
