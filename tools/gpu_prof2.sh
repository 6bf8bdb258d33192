#!/bin/bash
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_i4 -s 3 -c 1 -f -o gpurun_out/prof2_m16_nosplit python tools/prof_gemm.py 16 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_i4 -s 3 -c 1 -f -o gpurun_out/prof2_m4096_tall python tools/prof_gemm.py 4096 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep; cat gpurun_out/smi.txt 2>/dev/null
