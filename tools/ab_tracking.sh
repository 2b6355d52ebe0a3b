#!/bin/bash
# A/B of the tracking modes on the bench workload (device-resident value only). usage: tools/ab_tracking.sh <histories per exposure> [extra env...]
H=${1:-277778}
for t in 0 1; do
  echo "== DXMCB200_TRACKING=$t"
  DXMCB200_TRACKING=$t python bench.py --histories $H --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('value %.4g hist/s  ms/step %.1f  lookups/h %.2f  kernel ms/step %s' % (d['value'], d['ms_per_step'], r['lookups_per_history'], {k: round(v,1) for k,v in r['kernel_ms_per_step'].items()}))
"
done
