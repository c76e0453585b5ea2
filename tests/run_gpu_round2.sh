#!/bin/bash
# round-2 record session on one B200: parity suite, default bench + reference arm, detection benches, config-5 epoch, launch list
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 420 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
timeout 900 python bench.py --impl reference --steps ${REF_STEPS:-8} --warmup 2 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_ref_$TAG.err
for w in det640 det1280; do
timeout 900 python bench.py --workload $w --steps 8 --warmup 3 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "$w rc=$?"; tail -3 gpurun_out/bench_${w}_$TAG.err
done
timeout 900 python bench.py --epoch-samples 1000000 --warmup 5 > gpurun_out/bench_epoch_n1_$TAG.json 2> gpurun_out/bench_epoch_n1_$TAG.err; echo "epoch rc=$?"; tail -3 gpurun_out/bench_epoch_n1_$TAG.err
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python - <<PY
import json
for f in ("bench_$TAG","bench_ref_$TAG","bench_det640_$TAG","bench_det1280_$TAG","bench_epoch_n1_$TAG"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); e=d.get("e2e") or {}; r=d.get("roofline") or {}
        print(f, "value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"frac",r.get("frac"),"e2e",round(e.get("value",0),1),"cpu",(d.get("cpu_baseline") or {}).get("value"), d.get("epoch"))
    except Exception as ex: print(f,"parse failed",ex)
PY
