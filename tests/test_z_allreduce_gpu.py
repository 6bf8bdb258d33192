"""The push all-reduce kernel (csrc/comm_kernels.cuh) on ONE GPU: with world == 1 the "peer" table holds the local buffer,
so the full protocol runs (push into slot [e % 3][0], reset of the previous buffer to the sentinel, poll, FP32 reduction,
call counter) and the result must equal the input bit for bit (-0.0 arrives as +0.0) -- eagerly, with sizes that grow and
shrink between calls and leave CTAs without work, and across CUDA-graph replays (the buffer rotation is driven by the
device-side counter).  The N >= 2 behaviour is covered by tools/tp_check.py under torchrun
(profiles/r02_tp_check_n{2,8}.jsonl) and the host logic by tests/test_tp_gloo.py."""
import pytest
import torch

from atom_b200 import _lib

pytestmark = pytest.mark.gpu


def _setup(slot, dev):
    buf = torch.full((3 * 1 * slot,), -32768, dtype=torch.int16, device=dev)  # [3][world=1][slot] of 0x8000 (FP16 -0.0)
    state = torch.zeros(_lib.lib().atom_allreduce_state_words(), dtype=torch.int32, device=dev)
    bufs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=dev)
    return buf, state, bufs


def _call(x, out, st, slot):
    buf, state, bufs = st
    _lib.check(_lib.lib().atom_allreduce_push_f16(x.data_ptr(), out.data_ptr(), bufs.data_ptr(), state.data_ptr(),
                                                  x.numel(), slot, 0, 1, torch.cuda.current_stream().cuda_stream), "allreduce_push_f16")


def test_world1_identity_and_epochs():
    dev = torch.device("cuda", 0)
    slot = 32 * 8192
    st = _setup(slot, dev)
    for it, n in enumerate([8, 64, 4096, 32 * 5120, slot, 2048, slot, 8, slot]):
        torch.manual_seed(it)
        x = (torch.randn(n, device=dev) * 5).half()
        x[::7] = -0.0                                           # the sentinel pattern as payload: must arrive as +0.0
        x[3::11] = 0.0
        out = torch.full_like(x, 7.0)
        _call(x, out, st, slot)
        torch.cuda.synchronize()
        assert torch.equal(out, x), f"call {it} (numel {n})"   # (-0.0 == +0.0 under torch.equal)
        assert not (out.view(torch.int16) == -32768).any(), "-0.0 must be delivered as +0.0"
    assert int(st[1][:64].max()) == 9 and int(st[1][:64].min()) == 9      # every CTA counted every call
    torch.cuda.synchronize()
    # after the last call only the buffer it used holds payload; the one before it has been reset
    b3 = st[0].view(3, slot)
    assert (b3[(9 + 2) % 3] == -32768).all()


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
