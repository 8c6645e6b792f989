#!/bin/bash
for fl in "" "-DB200_ZST_FN=__stcg -DB200_AST_FN=__stcg" "-DB200_ZST_FN=__stwt -DB200_AST_FN=__stwt" "-DB200_ZST_FN=__stcs -DB200_AST_FN=__stwt"; do
  touch dspsr_b200/csrc/fastpath.cu; make -C dspsr_b200/csrc -j8 EXTRA="$fl" >/dev/null 2>&1
  for rep in 1 2; do python bench.py --steps 6 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$fl] value %.0f' % d['value'], ' '.join('%s %.4f' % (k, v['ms_per_launch']) for k,v in d['kernels'].items()))"; done
done
