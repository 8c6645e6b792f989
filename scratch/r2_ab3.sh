#!/bin/bash
mkdir -p gpurun_out/r2k
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg1 or pipeline_small or golden" > gpurun_out/r2k/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2k/pytest.log
tail -3 gpurun_out/r2k/pytest.log
L=$PWD/dspsr_b200
bash scratch/ab.sh e1i:B200_LIB=$L/libb200dsp_e1i.so new e1i2:B200_LIB=$L/libb200dsp_e1i.so new2
