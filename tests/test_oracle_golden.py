"""CPU suite: the oracle (oracle/) against the fixtures produced by the reference's own code.

Golden sources (tests/golden/make_golden.py): run_cpu_reorder_fp16_i4 (test_Reorder.cu:41-112),
run_cpu_activate_fp16_i4 (test_activate.cu:41-113) and model/quant.py + qLinearLayer.py.
The reference's CPU goldens divide by the scale where its kernels multiply by the reciprocal,
so -- exactly like the reference's own checks (test_Reorder.cu:269-318) -- quantised values may
differ by 1 LSB; everything structural (packing, reorder, scale layout, scale values) is exact.
"""
import os

import numpy as np
import torch

from oracle import oracle as O
from oracle import fakequant as FQ


def _lsb_diff_int4(p, q):
    return np.abs(O.unpack_int4(p).astype(np.int32) - O.unpack_int4(q).astype(np.int32))


def test_scale_layout_matches_reference_formulas():
    # ops/__init__.py:137-138 and Reorder.cuh:39-50
    for m in list(range(1, 70)) + [127, 128, 129, 4096]:
        ref = m // 16 * 64 + 64 - (1 - (m % 16) // 8) * (8 - (m % 8)) * 8
        assert O.scale_size(m) == ref
        used = set()
        for r in range(m):
            si = O.scale_index(r)
            assert si == (r // 16) * 64 + (r % 8) * 8 + (r // 8) % 2
            for j in range(4):
                assert si + 2 * j < ref
                used.add(si + 2 * j)
        assert len(used) == 4 * m  # replicas never collide


def test_reorder_against_reference_cpu_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_cpu_reorder_21x4096.npz"))
    o8, o4, s8, s4 = O.reorder_fp16_i4(g["x"], g["idx"])
    assert np.array_equal(s8.view(np.uint16), g["s8"].view(np.uint16))      # scales: bit exact
    assert np.array_equal(s4.view(np.uint16), g["s4"].view(np.uint16))
    d8 = np.abs(o8.astype(np.int32) - g["o8"].astype(np.int32))
    d4 = _lsb_diff_int4(o4, g["o4"])
    assert d8.max() <= 1 and d4.max() <= 1
    assert (d8 != 0).mean() < 2e-3 and (d4 != 0).mean() < 2e-3             # only reciprocal-vs-divide ties


def test_activate_against_reference_cpu_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_cpu_activate_5x11008.npz"))
    o8, o4, s8, s4 = O.activate_fp16_i4(g["x"], g["x2"])
    assert np.array_equal(s8.view(np.uint16), g["s8"].view(np.uint16))
    assert np.array_equal(s4.view(np.uint16), g["s4"].view(np.uint16))
    assert np.abs(o8.astype(np.int32) - g["o8"].astype(np.int32)).max() <= 1
    assert _lsb_diff_int4(o4, g["o4"]).max() <= 1


def test_rmsnorm_against_float64_restatement():
    # the reference's own CPU golden for K4 is uncompilable (see make_golden.py); restate RMSNorm.cuh:100-151
    rng = np.random.default_rng(3)
    m, h = 9, 4096
    x = (rng.standard_normal((m, h)) * 1.5).astype(np.float16)
    w = (1 + 0.2 * rng.standard_normal(h)).astype(np.float16)
    idx = rng.permutation(h).astype(np.int16)
    o8, o4, s8, s4 = O.rmsnorm_fp16_i4(x, w, idx, 1e-5)
    xf = x.astype(np.float64)
    y = (xf * w.astype(np.float64) / np.sqrt((xf ** 2).mean(1, keepdims=True) + 1e-5)).astype(np.float16)
    r8, r4, rs8, rs4 = O.reorder_fp16_i4(y, idx)
    # fp16 rounding of the normalised value may flip near ties: allow 1 LSB and ~1 ulp on scales
    assert np.abs(o8.astype(np.int32) - r8.astype(np.int32)).max() <= 1
    assert _lsb_diff_int4(o4, r4).max() <= 1
    assert np.allclose(s4.astype(np.float32), rs4.astype(np.float32), rtol=2e-3, atol=0)
    assert np.allclose(s8.astype(np.float32), rs8.astype(np.float32), rtol=2e-3, atol=0)


def test_gemm_oracle_against_float64_dequant():
    """A.3: with pair-shared B scales the faithful pairing equals the plain one, and the fp16 result
    equals float64 dequant-GEMM up to the documented fp16 roundings (rs product + output)."""
    m, n, k = 19, 256, 512
    t = O.make_gemm_inputs(m, n, k, seed=11, pair_shared=True)
    d_f = O.gemm_i4_o16(*t, faithful=True)
    d_p = O.gemm_i4_o16(*t, faithful=False)
    assert np.array_equal(d_f.view(np.uint16), d_p.view(np.uint16))
    a, b, a_s, b_s, ak, bk, aks, bks = t
    A, B = O.unpack_int4(a).astype(np.float64), O.unpack_int4(b).astype(np.float64)
    sa = O.a_scale_from_layout(a_s, m).astype(np.float64)          # [G, M]
    sak = O.a_scale_from_layout(aks, m).astype(np.float64)
    ref, bound = np.zeros((m, n)), np.zeros((m, n))
    terms = [(A[:, g * 128:(g + 1) * 128] @ B[:, g * 128:(g + 1) * 128].T, sa[g], b_s[g].astype(np.float64))
             for g in range(k // 128 - 1)]
    terms.append((ak.astype(np.float64) @ bk.astype(np.float64).T, sak, bks.astype(np.float64)))
    for c, ra, cb in terms:
        rs = ra[:, None] * cb[None, :]
        ref += c * rs
        # one fp16 rounding of the scale product: half an ulp, where ulp >= 2^-24 (fp16 subnormals)
        bound += np.abs(c) * np.maximum(2.0 ** -11 * rs, 2.0 ** -25)
    err = np.abs(d_f.astype(np.float64) - ref)
    assert (err <= 1.01 * (bound + 2.0 ** -11 * np.abs(ref)) + 1e-6).all()   # + the fp16 output rounding


def test_gemm_oracle_faithful_pairing_quirk():
    """GEMM.cuh:413-431: rows with m%16<8 use sB[n&~1], the others sB[n|1]."""
    m, n, k = 16, 128, 256
    t = list(O.make_gemm_inputs(m, n, k, seed=5, pair_shared=False))
    d_f = O.gemm_i4_o16(*t, faithful=True)
    bs, bks = t[3].copy(), t[7].copy()
    lo = list(t); lo[3] = np.repeat(bs[:, 0::2], 2, axis=1); lo[7] = np.repeat(bks[0::2], 2)
    hi = list(t); hi[3] = np.repeat(bs[:, 1::2], 2, axis=1); hi[7] = np.repeat(bks[1::2], 2)
    d_lo, d_hi = O.gemm_i4_o16(*lo, faithful=False), O.gemm_i4_o16(*hi, faithful=False)
    assert np.array_equal(d_f[:8].view(np.uint16), d_lo[:8].view(np.uint16))
    assert np.array_equal(d_f[8:].view(np.uint16), d_hi[8:].view(np.uint16))


def test_gemm_rows_subset_matches_full():
    t = O.make_gemm_inputs(33, 128, 384, seed=2)
    full = O.gemm_i4_o16(*t)
    sub = O.gemm_i4_o16(*t, rows=[0, 7, 8, 32])
    assert np.array_equal(full[[0, 7, 8, 32]].view(np.uint16), sub.view(np.uint16))


def test_o4_epilogue_roundtrip():
    """A.4: q*scale - zero recovers the FP32 accumulator to within scale/2 when the data are positive
    (where the reference's abs() quirk is harmless); the faithful mode differs from signed mode otherwise."""
    t = O.make_gemm_inputs(8, 256, 384, seed=9)
    d16 = O.gemm_i4_o16(*t).astype(np.float32)
    q, ds = O.gemm_i4_o4(*t, signed_minmax=True)
    ds = ds.reshape(8, 2, 2).astype(np.float32)
    vals = np.stack([(q & 0xF), (q >> 4)], -1).reshape(8, 2, 128).astype(np.float32)
    deq = vals * ds[:, :, :1] - ds[:, :, 1:]
    assert np.abs(deq.reshape(8, 256) - d16).max() <= 0.51 * ds[:, :, 0].max() + 2e-2 * np.abs(d16).max()
    q2, ds2 = O.gemm_i4_o4(*t, signed_minmax=False)
    assert q2.shape == q.shape and ds2.shape == (8, 4)


def _kv_fixture(rng, B, H, P, L, lens):
    pages = sum((l + P - 1) // P for l in lens) + 2
    data = rng.integers(0, 256, (pages, L, 2, H, P, 64), dtype=np.uint8)
    param = np.stack([rng.uniform(0.01, 0.05, (pages, L, 2, H, P)), rng.uniform(0, 0.4, (pages, L, 2, H, P))], -1).astype(np.float16)
    perm = rng.permutation(pages)
    indptr, indices, last = [0], [], []
    c = 0
    for l in lens:
        npg = (l + P - 1) // P
        indices += list(perm[c:c + npg]); c += npg
        indptr.append(len(indices)); last.append((l - 1) % P + 1)
    return data, param, np.array(indptr, np.int32), np.array(indices, np.int32), np.array(last, np.int32)


def test_decode_oracle_against_torch_reference():
    """ref_batch_decode of tests/test_batch_decode_int4.py:41-74 restated: dequant, RoPE(q@len-1,k@0..), softmax."""
    rng = np.random.default_rng(0)
    B, H, P, L = 3, 2, 16, 2
    lens = [1, 37, 64]
    data, param, indptr, indices, last = _kv_fixture(rng, B, H, P, L, lens)
    q = rng.standard_normal((B, H, 128)).astype(np.float16)
    o = O.batch_decode_i4(q, data, param, indptr, indices, last, layer=1)
    inv = 1.0 / (10000 ** (np.arange(0, 128, 2) / 128.0))

    def rope(x, pos):
        f = np.concatenate([pos[:, None] * inv[None], pos[:, None] * inv[None]], -1)
        rot = np.concatenate([-x[..., 64:], x[..., :64]], -1)
        return x * np.cos(f) + rot * np.sin(f)

    for b in range(B):
        pg = indices[indptr[b]:indptr[b + 1]]
        for h in range(H):
            kd = data[pg, 1, 0, h].reshape(-1, 64)[:lens[b]]
            vd = data[pg, 1, 1, h].reshape(-1, 64)[:lens[b]]
            kp = param[pg, 1, 0, h].reshape(-1, 2)[:lens[b]].astype(np.float64)
            vp = param[pg, 1, 1, h].reshape(-1, 2)[:lens[b]].astype(np.float64)
            un = lambda d: np.stack([d & 0xF, d >> 4], -1).reshape(d.shape[0], 128).astype(np.float64)
            K = un(kd) * kp[:, :1] - kp[:, 1:]
            V = un(vd) * vp[:, :1] - vp[:, 1:]
            qq = rope(q[b, h].astype(np.float64)[None], np.array([lens[b] - 1.0]))[0]
            KK = rope(K, np.arange(lens[b], dtype=np.float64))
            s = KK @ qq / np.sqrt(128.0)
            p = np.exp(s - s.max()); p /= p.sum()
            ref = p @ V
            assert np.allclose(o[b, h].astype(np.float64), ref, rtol=2e-3, atol=2e-3)


def test_append_and_init_kv_oracle():
    rng = np.random.default_rng(1)
    B, H, P, L = 2, 4, 8, 2
    lens = [9, 16]
    data, param, indptr, indices, last = _kv_fixture(rng, B, H, P, L, lens)
    d0, p0 = data.copy(), param.copy()
    k = rng.integers(0, 256, (B, H, 64), dtype=np.uint8); v = rng.integers(0, 256, (B, H, 64), dtype=np.uint8)
    kp = rng.random((B, H, 2)).astype(np.float16); vp = rng.random((B, H, 2)).astype(np.float16)
    O.append_kv_i4(data, param, indptr, indices, last, k, v, kp, vp, layer=0)
    for b in range(B):
        page = indices[indptr[b] + (lens[b] - 1) // P]; e = (lens[b] - 1) % P
        assert np.array_equal(data[page, 0, 0, :, e], k[b]) and np.array_equal(data[page, 0, 1, :, e], v[b])
        assert np.array_equal(param[page, 0, 0, :, e], kp[b]) and np.array_equal(param[page, 0, 1, :, e], vp[b])
    changed = (data != d0).sum()
    assert changed <= 2 * B * H * 64
    # prefill append of whole sequences == token-by-token placement
    data2, param2 = d0.copy(), p0.copy()
    tot = sum(lens)
    K = rng.integers(0, 256, (tot, H, 64), dtype=np.uint8); V = rng.integers(0, 256, (tot, H, 64), dtype=np.uint8)
    KP = rng.random((tot, H, 2)).astype(np.float16); VP = rng.random((tot, H, 2)).astype(np.float16)
    sl = np.array([0, lens[0], tot], np.int32)
    O.init_kv_i4(data2, param2, indptr, indices, last, K, V, KP, VP, sl, layer=1)
    t = 0
    for b in range(B):
        for j in range(lens[b]):
            page = indices[indptr[b] + j // P]
            assert np.array_equal(data2[page, 1, 0, :, j % P], K[t]) and np.array_equal(param2[page, 1, 1, :, j % P], VP[t])
            t += 1
    assert np.array_equal(data2[:, 0], d0[:, 0])  # other layer untouched


def test_fakequant_port_against_reference_python(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_py_fakequant.npz"))
    args = FQ.w4a4_args()
    wq = FQ.fq_linear_weight(torch.from_numpy(g["w0"]), args)
    assert torch.equal(wq, torch.from_numpy(g["wq"]))
    xq = FQ.fq_activation(torch.from_numpy(g["x"]), args)
    assert torch.equal(xq, torch.from_numpy(g["xq"]))
    y = FQ.fq_linear_forward(torch.from_numpy(g["x"]), wq, args)
    assert torch.allclose(y, torch.from_numpy(g["y"]), rtol=1e-5, atol=1e-5)
    assert torch.equal(FQ.fq_tensor(torch.from_numpy(g["t"]), 4, 128, True, 0.9), torch.from_numpy(g["t_sym"]))
    assert torch.equal(FQ.fq_tensor(torch.from_numpy(g["t"]), 4, 128, False, 1.0), torch.from_numpy(g["t_asym"]))
    kq = FQ.fq_tensor(torch.from_numpy(g["kv"]).reshape(-1, 128), 4, 0, False, 1.0).reshape(g["kv"].shape)
    assert torch.equal(kq, torch.from_numpy(g["kq"]))
