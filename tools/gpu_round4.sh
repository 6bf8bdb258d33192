#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x -p no:cacheprovider 2>&1 | tail -8
timeout 300 python tools/layer_bench.py 2>&1 | tail -2 | tee gpurun_out/layer_7b.json
timeout 300 python tools/layer_bench.py --hidden 5120 --inter 13824 --heads 40 --batch 32 --kvlen 1024 --layers 40 2>&1 | tail -1 | tee gpurun_out/layer_13b.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/layer_launches.csv python tools/layer_bench.py --copies 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/layer_launches.csv')) if len(r)>10 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host, Kernel Name, ..., Metric Name, Unit, Value
agg={}
for r in rows[-16*2:]:
    name=r[4].split('(')[0][:70]; agg[name]=agg.get(name,0)+float(r[-1])
tot=sum(agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1]): print(f"{v/2/1000:8.2f} us/step {100*v/tot:5.1f}%  {k}")
PY
