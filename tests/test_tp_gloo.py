"""CPU suite: the tensor-parallel plumbing (slice arithmetic, column/row sharding, the single all-reduce per
column->row pair) on 2 gloo ranks.  Kernels cannot run here, so the GEMM inside RowParallelLinearInt4 is replaced
by the CPU oracle (test infrastructure); what is under test is atom_b200.tp itself."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from atom_b200 import tp


def test_split_sizes():
    assert tp.split_sizes(8192, 8) == [1024] * 8
    assert tp.split_sizes(22016, 8) == [2816] * 4 + [2688] * 4          # 21.5 groups -> 22 / 21
    assert sum(tp.split_sizes(11008, 4)) == 11008 and all(s % 128 == 0 for s in tp.split_sizes(11008, 4))
    assert tp.slice_range([2816, 2816, 2688], 2) == (5632, 8320)
    with pytest.raises(ValueError):
        tp.split_sizes(512, 4)           # 128-channel slices cannot hold a group and a keeper
    with pytest.raises(ValueError):
        tp.split_sizes(1000, 2)


def _oracle_gemm(a, b, a_s, b_s, ak, bk, aks, bks):
    from oracle import oracle as O
    n, g = b.shape[0], b_s.shape[0]
    bs = b_s.reshape(-1)[: g * n].view(g, n)
    d = O.gemm_i4_o16(a.numpy(), b.numpy(), a_s.numpy(), bs.numpy(), ak.numpy(), bk.numpy(), aks.numpy(), bks[:n].numpy())
    return torch.from_numpy(d.astype(np.float32))     # fp32 so that gloo can all-reduce it


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    m, k, n = 5, 768, 64
    layer = tp.RowParallelLinearInt4(k, n, rank, world, gemm_fn=_oracle_gemm)
    k0, k1 = layer.k0, layer.k1
    # every rank builds the same full-size random problem, then keeps its K-slice as a self-contained operand
    parts = []
    for r in range(world):
        kr = tp.split_sizes(k, world)[r]
        parts.append(O.make_gemm_inputs(m, n, kr, seed=100 + r))
    mine = parts[rank]
    with torch.no_grad():
        layer.local.weight_int4.copy_(torch.from_numpy(mine[1])); layer.local.weight_int8.copy_(torch.from_numpy(mine[5]))
        g = mine[3].shape[0]
        layer.local.scale_int4.reshape(-1)[: g * n].copy_(torch.from_numpy(mine[3]).reshape(-1))
        layer.local.scale_int8[:n].copy_(torch.from_numpy(mine[7]))
    y = layer((torch.from_numpy(mine[4]), torch.from_numpy(mine[0]), torch.from_numpy(mine[6]), torch.from_numpy(mine[2])))
    ref = sum(O.gemm_i4_o16(*p).astype(np.float32) for p in parts)
    ok = np.allclose(y.numpy(), ref, rtol=0, atol=0) and (k1 - k0) == tp.split_sizes(k, world)[rank]
    # column-parallel slices tile the output exactly
    col = tp.ColumnParallelLinearInt4(256, 512, "fp16", rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, (col.n0, col.n1))
    ok = ok and gathered == [(0, 256), (256, 512)]
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_row_parallel_allreduce_and_column_slices_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
