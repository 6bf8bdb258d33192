"""Collectives of the tensor-parallel W4A4 layers (BASELINE config #5).  NEW functionality -- the reference has no
multi-GPU code.

`PushAllReduce` is the latency-oriented all-reduce of the row-parallel projections: a hand-written one-shot kernel
(csrc/comm_kernels.cuh) that pushes every rank's FP16 partial into all peers' receive buffers over NVLink peer mappings
and reduces locally as the payload arrives (sentinel-filled rotating buffers: no flags, no fence, no barrier) -- one
launch, CUDA-graph capturable.  The peer mappings come from
torch.distributed._symmetric_memory (plumbing only: allocation + pointer exchange).  `make_allreduce` falls back to
ncclAllReduce (torch.distributed.all_reduce) when symmetric memory cannot be set up, and says which one it returned.
"""
import os

import torch
import torch.distributed as dist

from . import _lib


class NcclAllReduce:
    name = "nccl (torch.distributed.all_reduce)"

    def __init__(self, group=None):
        self.group = group

    def __call__(self, x):
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.group)
        return x


class PushAllReduce:
    name = "push (atom_b200 one-shot kernel over NVLink peer memory)"
    CTAS = 64
    fuse = True        # tp.py may fold the two halves into the row-parallel GEMM's epilogue and the following add+RMSNorm launch

    def __init__(self, max_numel, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.group, self.world, self.rank = group, dist.get_world_size(group), dist.get_rank(group)
        self.slot = (int(max_numel) + 7) // 8 * 8
        name = group.group_name
        if hasattr(symm_mem, "is_symm_mem_enabled_for_group") and not symm_mem.is_symm_mem_enabled_for_group(name):
            symm_mem.enable_symm_mem_for_group(name)
        self.buf = symm_mem.empty(3 * self.world * self.slot, dtype=torch.float16, device=device)
        self.buf.view(torch.int16).fill_(-32768)          # 0x8000 = FP16 -0.0: "not yet arrived"
        self.hbuf = symm_mem.rendezvous(self.buf, group)
        self.state = torch.zeros(_lib.lib().atom_allreduce_state_words(), dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                  # every rank's buffers are initialised and mapped before anybody pushes
        from .ops import ArHandle
        # for the fused entry points (ops.dense_layer_gemm_i4_fp16_push / reduce_add_rmsnorm_fp16_i4): same buffers, same counter
        self.handle = ArHandle(self.hbuf.buffer_ptrs_dev, self.state, self.slot, self.rank, self.world)

    def __call__(self, x):
        if x.dtype != torch.float16 or not x.is_contiguous() or x.numel() % 8 or x.numel() > self.slot:
            raise RuntimeError("PushAllReduce: contiguous float16 tensor with numel % 8 == 0 and numel <= the configured maximum")
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().atom_allreduce_push_f16(x.data_ptr(), out.data_ptr(), self.hbuf.buffer_ptrs_dev,
                                                          self.state.data_ptr(), x.numel(), self.slot,
                                                          self.rank, self.world, torch.cuda.current_stream(x.device).cuda_stream),
                       "allreduce_push_f16")
        return out


def make_allreduce(max_numel, device, group=None):
    """The all-reduce the tensor-parallel layers use: the push kernel unless ATOM_B200_TP_ALLREDUCE=nccl or symmetric memory is
    unavailable.  All ranks must take the same branch, so the outcome of the set-up is agreed on with one all-reduce."""
    if dist.get_world_size(group) == 1:
        return None
    want_push = os.environ.get("ATOM_B200_TP_ALLREDUCE", "push") != "nccl"
    ar, err = None, ""
    if want_push:
        try:
            ar = PushAllReduce(max_numel, device, group)
        except Exception as e:  # noqa: BLE001
            err = f" (push set-up failed: {str(e)[:160]})"
    ok = torch.tensor([1 if ar is not None else 0], device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if ok.item() == 1:
        return ar
    nccl = NcclAllReduce(group)
    nccl.name = NcclAllReduce.name + err
    return nccl
