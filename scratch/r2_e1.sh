#!/bin/bash
# K2 in-place split + bulk stores (tile-image Z), response prefetch, K1 TMA stores in chunked A': A/B on one box
mkdir -p gpurun_out/r2k
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/r2k/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2k/pytest.log
tail -3 gpurun_out/r2k/pytest.log
L=$PWD/dspsr_b200
bash scratch/ab.sh base:B200_LIB=$L/libb200dsp_base.so e1h:B200_LIB=$L/libb200dsp_e1h.so new base2:B200_LIB=$L/libb200dsp_base.so e1h2:B200_LIB=$L/libb200dsp_e1h.so new2
