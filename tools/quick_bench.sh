#!/bin/sh
# quick device-resident timing of a few samplers (development aid)
python bench.py --steps 30 --warmup 3 --no-cpu --no-detail --no-density --no-e2e "$@" 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%s: %.1f Gdraws/s  %.3f ms  hbm_frac %.3f' % (d['metric'], d['value'], d['ms_per_step'], d['roofline']['frac']))
for k,v in (d.get('detail') or {}).items(): print('  ', k, round(v['gdraws_per_s'],1), round(v['kernel_ms'],3), round(v['hbm_frac'],3))
"
