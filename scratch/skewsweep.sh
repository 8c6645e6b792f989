#!/bin/bash
python -m pytest tests -m gpu -x -q -k "cfg1" 2>&1 | tail -1
for sk in 0 1500 3000 5000; do B200_K1_SKEW=$sk python bench.py --steps 6 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('skew $sk value %.0f' % d['value'], ' '.join('%s %.4f' % (k, v['ms_per_launch']) for k,v in d['kernels'].items()))"; done
