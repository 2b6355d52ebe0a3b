#!/bin/bash
# sweep of runtime knobs on the bench workload (device-resident value only). usage: tools/sweep.sh <histories per exposure>
H=${1:-277778}
run() {
  echo "== $*"
  env "$@" python bench.py --histories $H --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('value %.4g hist/s  ms/step %.1f  lookups/h %.2f  kernel ms/step %s' % (d['value'], d['ms_per_step'], r['lookups_per_history'], {k: round(v,1) for k,v in r['kernel_ms_per_step'].items()}))
"
}
run DXMCB200_TRACKING=1
run DXMCB200_TRACKING=1 DXMCB200_BATCH=16,26
run DXMCB200_TRACKING=1 DXMCB200_BATCH=24,26
run DXMCB200_TRACKING=1 DXMCB200_INTERACT_FULL=1
run DXMCB200_TRACKING=1 DXMCB200_BATCH=16,26 DXMCB200_INTERACT_FULL=1
run DXMCB200_TRACKING=1 DXMCB200_BRICK_MM=32
run DXMCB200_TRACKING=1 DXMCB200_PIPES=3
