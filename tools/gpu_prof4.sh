#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:batch_decode -s 2 -c 1 -f -o gpurun_out/prof_decode3 python tools/layer_bench.py --copies 1 > /dev/null 2>&1
ls -la gpurun_out/prof_decode3.ncu-rep
