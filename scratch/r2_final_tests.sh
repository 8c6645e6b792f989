#!/bin/bash
# the whole GPU suite on a 2-GPU box: the multi-GPU tests run instead of skipping
mkdir -p gpurun_out/r2u
nvidia-smi -L > gpurun_out/r2u/gpus.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2u/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u/pytest.log
tail -4 gpurun_out/r2u/pytest.log
