#!/bin/bash
# One GPU session: parity suite, bench, launch list, full ncu capture of the dominant kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/rc.txt
tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/rc.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
SMALL="python bench.py --steps 3 --warmup 3 --pool-cards 256 --pool-bgs 128 --no-e2e --no-cpu-baseline"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?" | tee -a gpurun_out/rc.txt
$SMALL > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_encoder -s 3 -c 2 -f -o gpurun_out/prof_encoder $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a gpurun_out/rc.txt
tail -3 gpurun_out/ncu_full.log
