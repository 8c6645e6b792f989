#!/bin/bash
# per-kernel ablations: B200_DBGn bit0 = no FFT, bit1 = no epilogue/stores, bit2 = no global loads.
# Needs a library built with the switches compiled in:  make -C dspsr_b200/csrc EXTRA=-DB200_ABLATION  (the product build
# compiles them out; the 32.16.16 plan of K3 has no switches)
run() { python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', ' '.join('%s %.4f' % (k, v['ms_per_launch']) for k,v in d['kernels'].items()))"; }
run base
for f in 1 2 4 3 5 6 7; do B200_DBG1=$f B200_DBG2=$f B200_DBG3=$f run dbg$f; done
