#!/bin/bash
# last single-GPU visit of a round: full GPU suite, a stress loop over the fused-op parity tests (flake hunt), smoke, both bench
# arms, layer records, per-kernel breakdown
OUT=gpurun_out; mkdir -p $OUT
bash tools/gpu_round.sh tests smoke
echo "== stress: fused / GEMM parity x6"
for i in 1 2 3 4 5 6; do
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_z_allreduce_gpu.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "gemm or fused or world1" > $OUT/pytest_stress_$i.txt 2>&1
  echo "run $i rc=$? $(tail -1 $OUT/pytest_stress_$i.txt)"; grep -h "AssertionError" $OUT/pytest_stress_$i.txt | head -3
done
bash tools/gpu_round.sh bench layer
bash tools/gpu_prof.sh layer
du -sh $OUT
