#!/bin/bash
run() { python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', ' '.join('%s %.4f' % (k, v['ms_per_launch']) for k,v in d['kernels'].items() if k=='cols_fwd'))"; }
run base
for f in 1 2 4 3 5 6 7; do B200_DBG1=$f run dbg1=$f; done
