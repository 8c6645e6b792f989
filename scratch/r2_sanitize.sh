#!/bin/bash
# compute-sanitizer over the round-2 kernels: k2_g2 (in-place split, bulk stores), K3 tile-image loads, k_conv64k (clusters)
out=gpurun_out/r2o; mkdir -p $out
K='test_pipeline_cfg1 or test_pipeline_small or test_cluster_convolution_kernel or test_pipeline_cfg3_meerkat'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > $out/memcheck.log 2>&1; echo "rc=$?" >> $out/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > $out/racecheck.log 2>&1; echo "rc=$?" >> $out/racecheck.log
tail -6 $out/memcheck.log; tail -6 $out/racecheck.log
