#!/bin/bash
# full ncu capture of the encoder pixel kernels on a small bench configuration (after a plain run exits 0)
set -u
mkdir -p gpurun_out
SMALL="python bench.py --steps 3 --warmup 3 --pool-cards 256 --pool-bgs 128 --no-e2e --no-cpu-baseline"
KERN="${1:-k_encoder|k_background|k_foreground}"
$SMALL > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$KERN" -s ${2:-9} -c ${3:-3} -f -o gpurun_out/prof_pixels $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -2 gpurun_out/ncu_full.log
