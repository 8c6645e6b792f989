#!/bin/bash
for fl in "" "-DB200_K2_PAIRED_TILES" "-DB200_K2_PAIRED_TILES -DB200_ZST_FN=__stcg" "-DB200_K2_PAIRED_TILES -DB200_ZST_FN=__stwt"; do
  touch dspsr_b200/csrc/fastpath.cu; make -C dspsr_b200/csrc -j8 EXTRA="$fl" >/dev/null 2>&1
  python -m pytest tests -m gpu -x -q -k "test_pipeline_cfg1 and not bench" 2>&1 | tail -1
  python bench.py --steps 6 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$fl] value %.0f' % d['value'], ' '.join('%s %.4f' % (k, v['ms_per_launch']) for k,v in d['kernels'].items()))"
done
