#!/bin/bash
# launch list + full ncu capture of the two pixel kernels on a small bench configuration
set -u
mkdir -p gpurun_out
SMALL="python bench.py --steps 3 --warmup 3 --pool-cards 256 --pool-bgs 128 --no-e2e --no-cpu-baseline"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
$SMALL > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_encoder|k_background|k_foreground" -s 9 -c 6 -f -o gpurun_out/prof_pixels $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -2 gpurun_out/ncu_full.log
