#!/bin/bash
# quick GPU iteration: parity suite + small bench (+ optional launch list)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
SMALL="python bench.py --steps 5 --warmup 3 --pool-cards 256 --pool-bgs 128 --no-e2e --no-cpu-baseline"
$SMALL > gpurun_out/plain.log 2> gpurun_out/plain.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/plain.log; tail -3 gpurun_out/plain.err
if [ "${1:-}" = "launches" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10 and r[0].isdigit()]
d=collections.defaultdict(list)
for r in rows: d[r[4].split('(')[0]+' grid='+r[8]].append(int(r[-1]))
for k,v in d.items(): print(k,len(v),round(sum(v)/len(v)/1e3,1),'us')
PY
fi
