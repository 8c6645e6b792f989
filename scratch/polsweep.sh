#!/bin/bash
for fl in "" "-DB200_NO_STREAMING"; do
  touch dspsr_b200/csrc/fastpath.cu; make -C dspsr_b200/csrc -j8 EXTRA="$fl" >/dev/null 2>&1
  for b in 1 2 3 4; do python bench.py --steps 6 --warmup 3 --no-cpu --batch $b --parts 24 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$fl] batch $b value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), ' '.join('%s %.4f' % (k, v['ms_per_launch']) for k,v in d['kernels'].items()))"; done
done
