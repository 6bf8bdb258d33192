#!/bin/bash
# multi-GPU visit (gpurun --gpus N): correctness of the push all-reduce + TP layer, then bench.py at N ranks (weak scaling over
# independent GEMM problems + the tensor-parallel decode-layer record), optionally the NCCL variant for comparison
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo_n$N.txt 2>&1
echo "== tp_check"
timeout 240 $RUN --master-port 29516 tools/tp_check.py 2> $OUT/tp_check_n$N.err | tee $OUT/tp_check_n$N.jsonl | cut -c1-300
echo "rc=${PIPESTATUS[0]}"; tail -3 $OUT/tp_check_n$N.err
echo "== bench.py --gpus $N"
timeout 420 $RUN --master-port 29517 bench.py --gpus $N --steps 1000 --warmup 5 2> $OUT/bench_n$N.err | tee $OUT/bench_n$N.json | cut -c1-1500
tail -2 $OUT/bench_n$N.err
echo "== bench.py --gpus $N, NCCL all-reduce in the TP record"
ATOM_B200_TP_ALLREDUCE=nccl timeout 420 $RUN --master-port 29518 bench.py --gpus $N --steps 300 --warmup 5 --no-sweep 2> $OUT/bench_n${N}_nccl.err | tee $OUT/bench_n${N}_nccl.json | cut -c1-1500
tail -2 $OUT/bench_n${N}_nccl.err
[ "$2" = "noref" ] && exit 0
echo "== reference arm under torchrun (rank 0 only)"
timeout 300 $RUN --master-port 29519 bench.py --impl reference --gpus $N --steps 20 --warmup 3 2> $OUT/bench_ref_n$N.err | tee $OUT/bench_ref_n$N.json | cut -c1-600
