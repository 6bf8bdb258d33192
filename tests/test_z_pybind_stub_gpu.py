"""GPU suite: the reference-side binding (integration/punica_ops_b200.cc, a drop-in for punica.ops._kernels built by
integration/Makefile) driven the way the reference's Python drives its own extension -- pre-allocated outputs, eight
functions, the input recipes of the reference's tests/test_int4.py (bs = 7, hidden 4096, uint8 codes in [16, 128)) -- and
compared with atom_b200.ops on the same tensors (both end in the same C ABI)."""
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "integration", "_kernels_b200.so")


@pytest.fixture(scope="module")
def K():
    if not os.path.exists(_SO):
        pytest.skip("integration/_kernels_b200.so not built (make -C integration)")
    spec = importlib.util.spec_from_file_location("_kernels_b200", _SO)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _quant_out(bs, hidden, dev):      # punica/ops/__init__.py:144-153
    from atom_b200.ops import scale_size
    return (torch.empty((bs, 128), dtype=torch.int8, device=dev), torch.empty((bs, (hidden - 128) // 2), dtype=torch.int8, device=dev),
            torch.empty((scale_size(bs),), dtype=torch.float16, device=dev),
            torch.empty((hidden // 128 - 1, scale_size(bs)), dtype=torch.float16, device=dev))


def _same_quant(a, b, bs):
    from oracle import oracle as O
    sel = torch.tensor([O.scale_index(r) + 2 * j for r in range(bs) for j in range(4)], device=a[0].device)
    return torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2][sel], b[2][sel]) and torch.equal(a[3][:, sel], b[3][:, sel])


@torch.inference_mode()
def test_stub_quantise_ops_match_the_python_mirror(K):
    from atom_b200 import ops
    dev, bs, hidden = torch.device("cuda:0"), 7, 4096
    torch.manual_seed(0xabcd)
    a = torch.randn((bs, hidden), dtype=torch.float16, device=dev)
    b = torch.randn((bs, hidden), dtype=torch.float16, device=dev)
    idx = torch.randperm(hidden, device=dev).to(torch.int16)
    w = torch.randn(hidden, dtype=torch.float16, device=dev)
    out = _quant_out(bs, hidden, dev); K.activate_fp16_i4(a, b, bs, *out)
    assert _same_quant(out, ops.activate_fp16_i4(a, b), bs)
    out = _quant_out(bs, hidden, dev); K.reorder_fp16_i4(a, idx, *out)
    assert _same_quant(out, ops.reorder_fp16_i4(a, idx), bs)
    out = _quant_out(bs, hidden, dev); K.rmsnorm_fp16_i4(a, w, 1e-5, idx, *out)
    assert _same_quant(out, ops.rmsnorm_fp16_i4(a, w, idx, 1e-5), bs)


@torch.inference_mode()
def test_stub_gemms_match_the_python_mirror(K):
    from atom_b200 import ops
    dev, bs, hidden, g = torch.device("cuda:0"), 7, 4096, 128
    torch.manual_seed(0xabcd + 1)
    a = torch.randint(16, 128, (bs, (hidden - g) // 2), dtype=torch.uint8, device=dev)
    b = torch.randint(16, 128, (hidden, (hidden - g) // 2), dtype=torch.uint8, device=dev)
    a_scale = torch.randn((hidden // g - 1, ops.scale_size(bs)), dtype=torch.float16, device=dev) * 0.05
    b_scale = torch.randn((hidden // g - 1, hidden), dtype=torch.float16, device=dev) * 0.05
    a_keeper = torch.randint(0, 255, (bs, g), dtype=torch.uint8, device=dev)
    b_keeper = torch.randint(0, 255, (hidden, g), dtype=torch.uint8, device=dev)
    a_keeper_scale = torch.randn((ops.scale_size(bs),), dtype=torch.float16, device=dev) * 0.01
    b_keeper_scale = torch.randn((hidden,), dtype=torch.float16, device=dev) * 0.01
    t = (a, b, a_scale, b_scale, a_keeper, b_keeper, a_keeper_scale, b_keeper_scale)
    d = torch.empty((bs, hidden), dtype=torch.float16, device=dev)                      # punica/ops/__init__.py:163
    K.dense_layer_gemm_i4_fp16(*t, d)
    assert torch.equal(d, ops.dense_layer_gemm_i4_fp16(*t))
    d4 = torch.empty((bs, hidden // 2), dtype=torch.uint8, device=dev)                  # :174-176
    ds = torch.empty((bs, hidden // g * 2), dtype=torch.float16, device=dev)
    K.dense_layer_gemm_i4_o4(*t, d4, ds)
    r4, rs = ops.dense_layer_gemm_i4_o4(*t)
    assert torch.equal(d4, r4) and torch.equal(ds, rs)


@torch.inference_mode()
def test_stub_kv_ops_match_the_python_mirror(K):
    import numpy as np
    from atom_b200 import ops
    from tests.test_gpu_parity import _KV, _kv_fixture
    rng = np.random.default_rng(77)
    B, H, P, L = 4, 8, 16, 2
    lens = [3, 16, 40, 100]
    fx = _kv_fixture(rng, B, H, P, L, lens)
    a, b = _KV(*fx), _KV(*fx)
    dev = a.data.device
    k = torch.randint(0, 256, (B, H, 64), dtype=torch.uint8, device=dev); v = torch.randint(0, 256, (B, H, 64), dtype=torch.uint8, device=dev)
    kp = torch.rand((B, H, 2), dtype=torch.float16, device=dev); vp = torch.rand((B, H, 2), dtype=torch.float16, device=dev)
    K.append_kv_i4(a.data, a.param, a.indptr, a.indicies, a.last_page_offset, k, v, kp, vp, 1)
    ops.append_kv_i4(b, k, v, kp, vp, 1)
    assert torch.equal(a.data, b.data) and torch.equal(a.param, b.param)
    q = torch.randn((B, H, 128), dtype=torch.float16, device=dev)
    o = torch.empty_like(q)
    K.batch_decode_i4(o, q, a.data, a.param, a.indptr, a.indicies, a.last_page_offset, 1)
    assert torch.equal(o, ops.batch_decode_i4(q, b, 1))
