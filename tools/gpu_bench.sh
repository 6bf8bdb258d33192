#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench ours"; timeout 600 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench_ours.json; tail -3 gpurun_out/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; grep -c gemm_i4 gpurun_out/bench_launches.csv
