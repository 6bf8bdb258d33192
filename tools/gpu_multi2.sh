#!/bin/bash
# 2-GPU development visit: single-GPU tests of the changed pieces first (cheap), then the multi-GPU checks
OUT=gpurun_out; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== pytest (all-reduce, fused ops, gemm parity)"
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 -p no:cacheprovider > $OUT/pytest.txt 2>&1; echo "rc=$?"; tail -8 $OUT/pytest.txt
echo "== layer decode step (1 GPU)"
timeout 300 python tools/layer_bench.py 2>&1 | tail -1 | tee $OUT/layer_7b.json | cut -c1-300
echo "== tp_check"
timeout 240 $RUN --master-port 29516 tools/tp_check.py 2> $OUT/tp_check_n2.err | tee $OUT/tp_check_n2.jsonl | cut -c1-300
echo "rc=${PIPESTATUS[0]}"; grep -v Warning $OUT/tp_check_n2.err | tail -5
echo "== bench.py --gpus 2"
timeout 420 $RUN --master-port 29517 bench.py --gpus 2 --steps 300 --warmup 5 --no-sweep 2> $OUT/bench_n2.err | tee $OUT/bench_n2.json | cut -c1-300
grep -v Warning $OUT/bench_n2.err | tail -3
