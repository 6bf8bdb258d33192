"""Tiny driver for ncu: python tools/prof_gemm.py M FLAGS [N K] -> 6 launches of the GEMM on rotating buffers."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from atom_b200 import ops
m, flags = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
k = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
base = O.make_gemm_inputs(m, n, k, seed=1)
sets = [[torch.from_numpy(x).cuda() for x in base] for _ in range(20 if m <= 256 else 6)]
for i in range(6):
    ops.dense_layer_gemm_i4_fp16(*sets[i % len(sets)], flags=flags)
torch.cuda.synchronize()
