"""The push all-reduce kernel (csrc/comm_kernels.cuh) on ONE GPU: with world == 1 the "peer" tables hold the local buffers,
so the full protocol runs (push into slot [parity][0], flag raise + wait, FP32 reduction, epoch advance) and the result must
equal the input bit for bit -- eagerly, with sizes that leave CTAs without work, and across CUDA-graph replays (the parity
double-buffering is driven by the device-side epoch).  The N >= 2 behaviour is covered by tools/tp_check.py under torchrun
(profiles/r02_tp_check_n{2,8}.jsonl) and the host logic by tests/test_tp_gloo.py."""
import pytest
import torch

from atom_b200 import _lib

pytestmark = pytest.mark.gpu


def _setup(slot, dev):
    buf = torch.zeros(2 * 1 * slot, dtype=torch.float16, device=dev)          # [parity][world=1][slot]
    flags = torch.zeros(2 * 64 * 1, dtype=torch.int32, device=dev)            # [parity][AR_CTAS][world]
    epoch = torch.zeros(64, dtype=torch.int32, device=dev)
    bufs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=dev)
    flgs = torch.tensor([flags.data_ptr()], dtype=torch.int64, device=dev)
    return buf, flags, epoch, bufs, flgs


def _call(x, out, st, slot):
    buf, flags, epoch, bufs, flgs = st
    _lib.check(_lib.lib().atom_allreduce_push_f16(x.data_ptr(), out.data_ptr(), bufs.data_ptr(), flgs.data_ptr(), epoch.data_ptr(),
                                                  x.numel(), slot, 0, 1, torch.cuda.current_stream().cuda_stream), "allreduce_push_f16")


def test_world1_identity_and_epochs():
    dev = torch.device("cuda", 0)
    slot = 32 * 8192
    st = _setup(slot, dev)
    for it, n in enumerate([8, 64, 4096, 32 * 5120, slot, 2048, slot]):
        torch.manual_seed(it)
        x = (torch.randn(n, device=dev) * 5).half()
        out = torch.empty_like(x)
        _call(x, out, st, slot)
        torch.cuda.synchronize()
        assert torch.equal(out, x), f"call {it} (numel {n})"
    assert int(st[2].max()) == 7 and int(st[2].min()) == 7      # every CTA advanced its epoch once per call


def test_world1_graph_replay():
    dev = torch.device("cuda", 0)
    slot = 16 * 4096
    st = _setup(slot, dev)
    x = torch.zeros(slot, dtype=torch.float16, device=dev)
    outs = [torch.empty_like(x) for _ in range(3)]
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        _call(x, outs[0], st, slot)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            _call(x, outs[0], st, slot)            # three chained all-reduces per replay: odd count, so the parity of the
            _call(outs[0], outs[1], st, slot)      # first call alternates between replays
            _call(outs[1], outs[2], st, slot)
        for rep in range(5):
            torch.manual_seed(100 + rep)
            x.copy_((torch.randn(slot, device=dev) * 3).half())
            g.replay()
            s.synchronize()
            assert torch.equal(outs[2], x), f"replay {rep}"


def test_argument_validation():
    dev = torch.device("cuda", 0)
    st = _setup(1024, dev)
    x = torch.zeros(1024 + 8, dtype=torch.float16, device=dev)
    with pytest.raises(RuntimeError):
        _call(x, torch.empty_like(x), st, 1024)      # numel > slot
    y = torch.zeros(12, dtype=torch.float16, device=dev)
    with pytest.raises(RuntimeError):
        _call(y, torch.empty_like(y), st, 1024)      # numel % 8 != 0
