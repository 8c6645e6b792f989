#!/bin/bash
# one-kernel convolution path (clusterconv.cu): parity of the 65536-point convolution tests, three times over (the
# inter-CTA barriers cannot be checked by racecheck), plus the DSMEM-cluster build once
mkdir -p gpurun_out/r2m
L=$PWD/dspsr_b200
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg3 or 65536 or convolution or cluster or one_kernel or cpp_engine" 2>&1 | tail -1; done
[ -f $L/libb200dsp_dsm.so ] && B200_LIB=$L/libb200dsp_dsm.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg3 or cluster" 2>&1 | tail -1
