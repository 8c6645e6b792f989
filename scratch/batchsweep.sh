#!/bin/bash
for b in 2 4 8 16 32; do python bench.py --steps 6 --warmup 3 --no-cpu --batch $b 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batch $b value %.0f' % d['value'], ' '.join('%s %.4f' % (k, v['ms_per_launch']) for k,v in d['kernels'].items()))"; done
