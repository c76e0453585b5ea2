#!/bin/bash
# round-2 GPU session: parity suite (per-test timeout), default bench (+ reference arm), optional ncu captures
# usage: run_gpu_r2.sh TAG [full|new|none] [ref] [prof KERNEL_REGEX]
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
if [ "${2:-full}" = "full" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 420 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_$TAG.log
elif [ "${2:-}" = "new" ]; then
timeout 900 python -m pytest ${NEW_TESTS:-tests/test_gpu_rng.py} -x -q --timeout 420 > gpurun_out/pytest_new_$TAG.log 2>&1; echo "pytest new rc=$?"
tail -25 gpurun_out/pytest_new_$TAG.log
fi
timeout 900 python bench.py --steps ${STEPS:-20} --warmup 5 ${BENCH_ARGS:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$TAG.json"))
    e=d.get("e2e") or {}
    print("value",round(d["value"]),"ms/step",round(d["ms_per_step"],3),"kernel_ms",round(d["roofline"]["kernel_ms"],3),"frac",round(d["roofline"]["frac"],4),
          "e2e",round(e.get("value",0)),"raw",round((e.get("from_raw_arrays") or {}).get("value",0)),"bytes",round((e.get("from_bytes_objects") or {}).get("value",0)),
          "resident",round((e.get("resident_pool") or {}).get("value",0)),"cpu",(d.get("cpu_baseline") or {}).get("value"),"failed",d["stats"]["failed_samples"])
except Exception as ex: print("bench parse failed",ex)
PY
tail -5 gpurun_out/bench_$TAG.err
if [ "${3:-}" = "ref" ]; then
timeout 900 python bench.py --impl reference --steps ${REF_STEPS:-3} --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
cut -c1-200 gpurun_out/bench_ref_$TAG.json; tail -5 gpurun_out/bench_ref_$TAG.err
fi
if [ "${4:-}" = "prof" ]; then
SMALL="python bench.py --steps 3 --warmup 3 --pool-cards 256 --pool-bgs 128 --no-e2e --no-cpu-baseline"
$SMALL > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${5:-k_encoder|k_foreground}" -s ${6:-6} -c ${7:-2} -f -o gpurun_out/prof_$TAG $SMALL > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_$TAG.log
fi
