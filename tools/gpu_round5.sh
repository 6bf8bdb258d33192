#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_i4|quant_kernel|decode_kernel|append_kv" --csv --log-file gpurun_out/layer_launches.csv python tools/layer_bench.py --copies 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/layer_launches.csv')) if len(r)>10 and r[0].isdigit()]
n=len(rows); per=16  # launches per layer step
agg={}; order=[]
last=rows[-per:]
for r in last:
    name=r[4].replace('atom::','').split('(')[0][:60]; v=float(r[-1])
    order.append((name,v))
tot=sum(v for _,v in order)
for name,v in order: print(f"{v/1000:8.2f} us {100*v/tot:5.1f}%  {name}")
print("total us", tot/1000, "launches captured", n)
PY
