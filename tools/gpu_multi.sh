#!/bin/bash
# multi-GPU visit: bench.py at N ranks (weak scaling over independent GEMM problems) and the tensor-parallel layer
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 1000 --warmup 5 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | cut -c1-700
tail -2 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 tools/tp_bench.py 2> gpurun_out/tp_n$N.err | tee gpurun_out/tp_n$N.json
tail -2 gpurun_out/tp_n$N.err
if [ "$N" = "2" ]; then python tools/tp_bench.py --heads 64 2>/dev/null | tee gpurun_out/tp_n1.json; fi
