#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tools/layer_bench.py 2>&1 | tail -1 | tee gpurun_out/layer_7b.json
bash tools/gpu_round5.sh
