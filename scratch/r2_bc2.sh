#!/bin/bash
for i in 1 2 3; do timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -k "long_conv or cfg4" 2>&1 | grep -v "^$" | tail -6; done
