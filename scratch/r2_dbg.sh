#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cluster" 2>&1 | grep -E "passed|failed|Error|assert|^E " | head -30
