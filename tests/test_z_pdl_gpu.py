"""GPU suite: programmatic dependent launch.  The GEMM kernels are always launched with programmatic stream serialization
(they fetch only weights before griddepcontrol.wait); atom_set_pdl(1) / ATOM_B200_PDL=1 extends it to the quantise and KV
kernels (they then signal launch_dependents at once, so the GEMM behind them streams its weights while they run).  PDL must
not change a single bit: every chain is run with it off first and on second (eagerly and from a CUDA graph)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def pdl():
    from atom_b200 import _lib

    def set_(on):
        torch.cuda.synchronize()
        _lib.check(_lib.lib().atom_set_pdl(int(on)), "atom_set_pdl")
    yield set_
    set_(False)


def _chain(steps=12):
    """x -> (reorder+quantise -> W4A4 GEMM) x steps: every kernel consumes its predecessor's output."""
    from atom_b200 import ops, synth
    dev = torch.device("cuda:0")
    w = synth.gemm_operands(16, 512, 512, dev, seed=3)
    idx = torch.randperm(512, generator=torch.Generator().manual_seed(0)).to(torch.int16).to(dev)
    x = torch.randn(16, 512, generator=torch.Generator().manual_seed(1)).half().to(dev)
    for _ in range(steps):
        o8, o4, s8, s4 = ops.reorder_fp16_i4(x, idx)
        x = ops.dense_layer_gemm_i4_fp16(o4.view(torch.uint8), w[1], s4, w[3], o8, w[5], s8, w[7])
        x = x / x.float().abs().amax().clamp(min=1e-3).half()          # keep the chain finite: NaNs would compare unequal to themselves
    assert torch.isfinite(x).all()
    return x


@pytest.mark.timeout(300)
def test_pdl_kernel_chain_is_bit_identical(pdl):
    pdl(False)
    ref = _chain().clone()
    pdl(True)
    for _ in range(3):
        assert torch.equal(_chain(), ref)


@pytest.mark.timeout(300)
def test_pdl_decoder_layer_eager_and_graph(pdl):
    from atom_b200.cat_tensor import BatchLenInfo
    from atom_b200.kvcache import BatchedKvCacheInt4, KvCacheInt4, KvPoolInt4
    from atom_b200.llama import LlamaConfig, LlamaDecoderLayer
    dev = torch.device("cuda:0")
    cfg = LlamaConfig(hidden_size=512, intermediate_size=1024, num_attention_heads=4, num_hidden_layers=1, vocab_size=128)
    layer = LlamaDecoderLayer(cfg, 0).to(dev).init_random(5)
    pool = KvPoolInt4(1, 4, 128, capacity=32, block_len=16, device=dev)
    pool.buf.random_(0, 256); pool.param.uniform_(0.01, 0.05)
    caches = [KvCacheInt4(pool, n) for n in (33, 7, 100)]
    for c in caches:
        c.acquire_one()
    kv = BatchedKvCacheInt4(caches)
    blen = BatchLenInfo([], 3, dev)
    x = torch.randn(3, 512, device=dev, dtype=torch.float16)
    pdl(False)
    ref = layer(x, blen, None, kv).clone()
    pdl(True)
    assert torch.equal(layer(x, blen, None, kv), ref)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        layer(x, blen, None, kv)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            y = layer(x, blen, None, kv)
        for _ in range(3):
            g.replay()
        st.synchronize()
    assert torch.equal(y, ref)
