#!/bin/bash
# dev: throughput of the classic uniform-kernel variants (see ebm_launch_classic_uniform)
for v in ${VARIANTS:-0 1 2 3 4 5 6}; do
  EBM_CLASSIC_VARIANT=$v python bench.py --members ${MEMBERS:-18944} --years ${YEARS:-5} --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('variant $v', 'my/s %.0f' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'kernel_ms %.1f' % d['roofline']['kernel_ms'], 'nan', d['nan_flags'])
    elif 'rror' in l: print(l.strip())
"
done
