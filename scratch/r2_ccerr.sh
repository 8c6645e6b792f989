#!/bin/bash
L=$PWD/dspsr_b200
B200_LIB=$L/libb200dsp_e1i.so timeout 300 python scratch/cc_err.py 2>&1 | tail -5
timeout 300 python scratch/cc_err.py 2>&1 | tail -5
