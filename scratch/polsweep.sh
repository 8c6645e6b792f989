#!/bin/bash
for fl in "-DB200_K2_UNROLL=1" "-DB200_K2_UNROLL=2" "-DB200_K2_UNROLL=4"; do
  touch dspsr_b200/csrc/fastpath.cu; make -C dspsr_b200/csrc -j8 EXTRA="$fl" >/dev/null 2>&1
  grep -A2 "k2_c2ILj2048ELj1024ELb1" dspsr_b200/_build/fastpath.ptxas.log | grep -E "spill" | head -1
  python bench.py --steps 6 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$fl] value %.0f' % d['value'], ' '.join('%s %.4f' % (k, v['ms_per_launch']) for k,v in d['kernels'].items()))"
done
