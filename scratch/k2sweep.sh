#!/bin/bash
run() { python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', ' '.join('%s %.4f' % (k, v['ms_per_launch']) for k,v in d['kernels'].items()))"; }
run base
for f in 5 13 21 29; do B200_DBG2=$f run dbg2=$f; done
