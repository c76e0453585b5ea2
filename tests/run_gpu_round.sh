#!/bin/bash
# One GPU session for the record: parity suite, default bench, launch list, full ncu capture of the
# encoder pixel kernels at the bench's pool sizes (each ncu pass only after the same command exited 0 plain).
set -u
mkdir -p gpurun_out
TAG=${1:-v8}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_encoder|k_background|k_foreground" -s 9 -c 3 -f -o gpurun_out/prof_pixels_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -2 gpurun_out/ncu_full.log
