#!/bin/bash
# ncu captures of the GEMM kernels (launch durations + one full capture each)
set -u
mkdir -p gpurun_out
for cfg in "16 0" "16 1" "64 0" "128 2" "4096 2"; do
  set -- $cfg
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:gemm_i4 -c 6 --csv \
      --log-file gpurun_out/launch_m$1_f$2.csv python tools/prof_gemm.py $1 $2 > /dev/null 2>&1
  echo "== M=$1 flags=$2"; grep -E "gpu__time_duration|dram__bytes_read" gpurun_out/launch_m$1_f$2.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -6
done
ncu --set full --clock-control none --import-source on -k regex:gemm_i4 -s 3 -c 1 -f -o gpurun_out/prof_m16_split python tools/prof_gemm.py 16 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_i4 -s 3 -c 1 -f -o gpurun_out/prof_m4096_tall python tools/prof_gemm.py 4096 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
