#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --workload ${1:-det640} --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_det.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_det_pixels|k_det_place" -s ${2:-4} -c ${3:-3} -f -o gpurun_out/prof_det $CMD > gpurun_out/ncu_det_full.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/ncu_det_full.log
