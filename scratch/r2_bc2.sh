#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "long_convolution or convolution_voltages" 2>&1 | grep -v "^$" | tail -30
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -q --tb=short -k "long_" 2>&1 | grep -v "^$" | tail -12
