#!/bin/bash
# compute-sanitizer over the kernels added after scratch/r2_sanitize.sh: k_bc_* (long transforms), k_conv64k in group mode
# (counter barriers, L2 exchange matrices), k_chan_inv_small, the fused two-bit column pass, chunked k_expand_bins
out=gpurun_out/r2t; mkdir -p $out
K='test_long_convolution_kernels and (262144 or 524288) or test_long_convolution_detected or test_short_channel or test_one_kernel_convolution_detected or test_pipeline_twobit_fold or test_cluster_convolution_kernel and uwb or test_convolution_voltages and 262144'
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > $out/memcheck.log 2>&1; echo "rc=$?" >> $out/memcheck.log
timeout 700 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > $out/racecheck.log 2>&1; echo "rc=$?" >> $out/racecheck.log
tail -6 $out/memcheck.log; tail -6 $out/racecheck.log
