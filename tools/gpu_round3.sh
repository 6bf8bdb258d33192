#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "16 0" "16 1" "4096 2"; do timeout 120 python tools/gpu_check.py trace $cfg 2>&1 | tail -1 | tee -a gpurun_out/trace.jsonl | cut -c1-2500; done
