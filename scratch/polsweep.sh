#!/bin/bash
for fl in "" "-DB200_FOLD_DIRECT_ATOMICS"; do
  touch dspsr_b200/csrc/fastpath.cu; make -C dspsr_b200/csrc -j8 EXTRA="$fl" >/dev/null 2>&1
  python -m pytest tests -m gpu -x -q -k "pipeline_cfg1" 2>&1 | tail -1
  python bench.py --steps 6 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$fl] value %.0f' % d['value'], ' '.join('%s %.4f' % (k, v['ms_per_launch']) for k,v in d['kernels'].items()))"
done
