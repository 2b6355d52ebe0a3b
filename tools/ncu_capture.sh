#!/bin/bash
# ncu captures of the wave kernels as the two-pipeline bench launches them (4e8 histories per step). usage: tools/ncu_capture.sh <tag> [kernels...]
TAG=${1:-r2}; shift
KERNELS=${@:-"generateKernel transportKernel airWalkKernel interactKernel"}
CMD="python bench.py --histories 111112 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > /dev/null 2>&1
for k in $KERNELS; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o gpurun_out/${TAG}_$k $CMD > gpurun_out/${TAG}_$k.log 2>&1
  echo "$k: $(ls -la gpurun_out/${TAG}_$k.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
done
