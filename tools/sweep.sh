#!/bin/bash
# a few settings on the bench workload (device-resident value only).
# usage: tools/sweep.sh <histories per exposure> "[LIB=<variant>] ENV=.. ENV=..;ENV=..;..."   (LIB: dxmclib_b200/variants/<variant>.so replaces the library for that run)
H=${1:-277778}
cp dxmclib_b200/libdxmcb200.so /tmp/libdxmcb200.main.so
IFS=';' read -ra SETS <<< "${2:-DXMCB200_TRACKING=1}"
for s in "${SETS[@]}"; do
  echo "== $s"
  cp /tmp/libdxmcb200.main.so dxmclib_b200/libdxmcb200.so
  envs=""
  for w in $s; do
    case $w in
      LIB=*) cp dxmclib_b200/variants/${w#LIB=}.so dxmclib_b200/libdxmcb200.so ;;
      *) envs="$envs $w" ;;
    esac
  done
  env $envs python bench.py --histories $H --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('value %.4g hist/s  ms/step %.1f  lookups/h %.2f bricks/h %.2f  kernel ms/step %s' % (d['value'], d['ms_per_step'], r['lookups_per_history'], r.get('bricks_crossed_per_history', 0), {k: round(v,1) for k,v in r['kernel_ms_per_step'].items()}))
"
done
cp /tmp/libdxmcb200.main.so dxmclib_b200/libdxmcb200.so
