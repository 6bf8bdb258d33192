#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
echo "== pytest (gemm only)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=240 -x -p no:cacheprovider -k "gemm" > $OUT/pytest.txt 2>&1; echo "rc=$?"; tail -5 $OUT/pytest.txt
echo "== trace"; for cfg in "16 0" "16 1" "4096 2"; do timeout 120 python tools/gpu_check.py trace $cfg 2>&1 | tail -1 | tee -a $OUT/trace.jsonl | cut -c1-3000; done
echo "== gtime"; timeout 600 python tools/gpu_check.py gtime 2>$OUT/gtime.err | tee $OUT/gtime.jsonl; tail -3 $OUT/gtime.err
