#!/bin/bash
# cluster convolution kernel (clusterconv.cu): parity of the 65536-point convolution tests
mkdir -p gpurun_out/r2m
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg3 or 65536 or convolution or cluster" > gpurun_out/r2m/pytest_cc.log 2>&1; echo "rc=$?" >> gpurun_out/r2m/pytest_cc.log
tail -25 gpurun_out/r2m/pytest_cc.log
