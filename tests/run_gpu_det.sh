#!/bin/bash
# detection workloads: bench lines (configs[2], configs[3]) + launch list
set -u
mkdir -p gpurun_out
python bench.py --workload det640 --steps 5 --warmup 3 > gpurun_out/bench_det640.json 2> gpurun_out/det.err; echo "rc=$?"; cut -c1-420 gpurun_out/bench_det640.json
python bench.py --workload det1280 --steps 3 --warmup 2 > gpurun_out/bench_det1280.json 2>> gpurun_out/det.err; echo "rc=$?"; cut -c1-420 gpurun_out/bench_det1280.json
tail -3 gpurun_out/det.err
for W in det640 det1280; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_det.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_$W.csv')) if len(r)>10 and r[0].isdigit()]
d=collections.defaultdict(list)
for r in rows:
    if 'mtgv' in r[4] or r[4].startswith('k_'): d[r[4].split('(')[0]+' grid='+r[8]+' blk='+r[7]].append(int(r[-1]))
print('$W')
for k,v in d.items(): print(' ',k,len(v),round(sum(v)/len(v)/1e3,1),'us')
PY
done
