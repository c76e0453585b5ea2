#!/bin/bash
# multi-GPU session (gpurun --gpus N): default bench and the config-5 epoch, launched like the driver does
set -u
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r02}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench N=$N rc=$?"; tail -3 gpurun_out/bench_n${N}_$TAG.err
timeout 900 $RUN bench.py --gpus $N --epoch-samples 1000000 --warmup 5 > gpurun_out/bench_epoch_n${N}_$TAG.json 2> gpurun_out/bench_epoch_n${N}_$TAG.err; echo "epoch N=$N rc=$?"; tail -3 gpurun_out/bench_epoch_n${N}_$TAG.err
python - <<PY
import json
for f in ("bench_n${N}_$TAG","bench_epoch_n${N}_$TAG"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); e=d.get("e2e") or {}
        print(f,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(e.get("value",0)),"raw",round((e.get("from_raw_arrays") or {}).get("value",0)),"resident",round((e.get("resident_pool") or {}).get("value",0)),d.get("epoch"))
    except Exception as ex: print(f,"parse failed",ex)
PY
