#!/bin/bash
# One GPU visit: environment, micro-benchmarks, GEMM probes, the gpu test suite, timing.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.txt 2>&1
echo "== microbench" ; timeout 120 ./tools/microbench > $OUT/microbench.jsonl 2>&1; echo "rc=$?"; tail -40 $OUT/microbench.jsonl
echo "== diag" ; timeout 300 python tools/gpu_check.py diag > $OUT/diag.jsonl 2> $OUT/diag.err; echo "rc=$?"; cat $OUT/diag.jsonl | cut -c1-400; tail -5 $OUT/diag.err
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -q --timeout=240 -x -p no:cacheprovider > $OUT/pytest.txt 2>&1; echo "rc=$?"; tail -40 $OUT/pytest.txt
echo "== timing" ; timeout 600 python tools/gpu_check.py time > $OUT/timing.jsonl 2> $OUT/timing.err; echo "rc=$?"; cat $OUT/timing.jsonl; tail -5 $OUT/timing.err
