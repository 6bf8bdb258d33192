"""CPU suite: simulated layer (model/ surface, FP tensors on INT grids) -> exported real-INT4 operands -> the ORACLE's
restatement of the kernels.  This pins the whole export mapping without a GPU: reorder indices (norm / attention output /
down_proj folded into gate+up rows), nibble order, keeper split, flat [group][out] weight-scale addressing and the
ldmatrix-replicated activation-scale layout.

Method: every stage is compared on IDENTICAL inputs.  A quantiser stage (oracle kernel vs simulator Quantizer) may differ
by fp16 storage of the scale and by tie-breaking (roundf vs round-half-even): at most one grid step on a few elements.
A GEMM stage is checked twice: the oracle GEMM on the exported operands against a float64 product of the dequantised
operands (layout / addressing, ~1e-3 from the fp16 scale product and output), and the dequantised operands against the
simulator's fake-quantised weight.  End to end (quantisation noise compounding through
three stages) the two worlds agree to a few per cent, which is asserted last."""
import types

import numpy as np
import torch

from atom_b200 import modelutils
from atom_b200.export import int4_decoder_layer
from atom_b200.qllama import ToyLlamaDecoderLayer
from oracle import oracle as O


def _args():
    return types.SimpleNamespace(keep_fp_for_export=True, wbits=4, abits=4, w_sym=True, a_sym=True, weight_group_size=128, act_group_size=128,
                                 weight_channel_group=2, w_clip_ratio=0.85, a_clip_ratio=1.0, keeper=128, keeper_precision=3,
                                 exponential=False, tiling=0, quant_type="int", static=False, kv_clip_ratio=1.0, reorder=True,
                                 kv_cache=True)


def _np(t):
    return t.detach().cpu().numpy()


def _gemm(inp, lin, **kw):
    o8, o4, s8, s4 = inp
    g, n = lin.in_features // 128 - 1, lin.out_features
    b_scale = _np(lin.scale_int4).reshape(-1)[: g * n].reshape(g, n)
    args = (o4, _np(lin.weight_int4), s4, b_scale, o8, _np(lin.weight_int8), s8, _np(lin.scale_int8)[:n])
    return O.gemm_i4_o4(*args, **kw) if lin.out_dtype == "int4" else O.gemm_i4_o16(*args)


def _dequant(inp):
    """(o8, o4, s8, s4) of a quantise kernel -> float [M, hidden] in the kernel's (reordered) channel order."""
    o8, o4, s8, s4 = inp
    m = o8.shape[0]
    s4r = O.a_scale_from_layout(s4, m).astype(np.float32)
    s8r = O.a_scale_from_layout(s8, m).astype(np.float32)
    body = O.unpack_int4(o4).astype(np.float32).reshape(m, -1, 128) * s4r.T[:, :, None]
    return np.concatenate([body.reshape(m, -1), o8.astype(np.float32) * s8r[:, None]], 1), s4r, s8r


def _outlier_last(n, outliers, gen):
    """A calibration-style index: random order, the given outlier channels moved into the last 128 (the keeper)."""
    rest = [c for c in torch.randperm(n, generator=gen).tolist() if c not in outliers]
    return torch.tensor(rest + list(outliers))


OUTLIERS = (3, 100)


def _build(hidden=256, inter=512, heads=2, seed=0):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    a = _args()
    layers = [ToyLlamaDecoderLayer(hidden, inter, heads)]
    qkv, gu = _outlier_last(hidden, OUTLIERS, gen), _outlier_last(hidden, OUTLIERS, gen)
    idx = {f"layers.0.self_attn.{p}.input": qkv for p in ("q_proj", "k_proj", "v_proj")}
    idx.update({f"layers.0.mlp.{p}.input": gu for p in ("gate_proj", "up_proj")})
    idx["layers.0.self_attn.o_proj.input"] = torch.randperm(hidden, generator=gen)
    idx["layers.0.mlp.down_proj.input"] = torch.randperm(inter, generator=gen)
    modelutils.reorder_model_llama(layers, a, idx)
    modelutils.quantize_model_llama(layers, a)
    modelutils.add_act_quant_wrapper_llama(layers, a)
    return layers[0], a


def _input(rows, hidden, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, hidden, generator=g) * 0.5
    for c in OUTLIERS:
        x[:, c] *= 12
    return x.half()


def _check_quant_stage(kernel_out, sim, what):
    deq, s4r, s8r = _dequant(kernel_out)
    g = s4r.shape[0]
    step = np.concatenate([np.repeat(s4r.T, 128, axis=1), np.repeat(s8r[:, None], 128, axis=1)], 1)
    d = np.abs(deq - sim)
    assert (d <= 1.01 * step + 1e-6).all(), what             # never more than one grid step
    assert (d > 0.5 * step).mean() < 0.01, what              # and a real flip only on ties (< 1 % of the elements)
    return deq


def _dequant_weight(lin):
    """LinearInt4 operands -> float [out, in] (reordered input order), plus the per-row keeper step."""
    g, n = lin.in_features // 128 - 1, lin.out_features
    s4 = _np(lin.scale_int4).reshape(-1)[: g * n].reshape(g, n).astype(np.float32)
    s8 = _np(lin.scale_int8)[:n].astype(np.float32)
    body = O.unpack_int4(_np(lin.weight_int4)).astype(np.float32).reshape(n, g, 128) * s4.T[:, :, None]
    return np.concatenate([body.reshape(n, -1), _np(lin.weight_int8).astype(np.float32) * s8[:, None]], 1), s8


def _check_gemm_stage(out, x_dq, lin, qlin, what, tol=1e-2):
    """(1) the oracle GEMM on the exported operands == x_dq @ dequant(operands)^T: addressing / layout / nibble order
    (tolerance: the kernel multiplies the two scales in FP16 -- toy-sized products fall into fp16 subnormals, ~0.5 %);
    (2) dequant(operands) == the simulator's fake-quantised weight: body to fp16-scale rounding, keeper to one INT8
    step of the pair-shared scale (the simulator keeps one keeper scale per row, see QLinearLayer.int4_operands)."""
    wd, s8 = _dequant_weight(lin)
    ref = x_dq.astype(np.float64) @ wd.astype(np.float64).T
    assert np.abs(out.astype(np.float64) - ref).max() <= tol * np.abs(ref).max(), (what, np.abs(out - ref).max(), np.abs(ref).max())
    wq = _np(qlin.weight)
    assert np.abs(wd[:, :-128] - wq[:, :-128]).max() <= 1e-3 * np.abs(wq).max(), what
    assert (np.abs(wd[:, -128:] - wq[:, -128:]) <= 1.05 * s8[:, None] + 1e-6).all(), what   # two INT8 roundings, each <= step/2 (+ fp16 scale storage)


def test_exported_mlp_branch_stage_by_stage():
    q, a = _build()
    real = int4_decoder_layer(q, device=None)
    x = _input(7, 256, 1)
    n = real.post_attention_layernorm
    assert not any(getattr(m, "packed", False) for m in q.modules())      # exporting leaves the simulator untouched
    h = O.rmsnorm_fp16_i4(_np(x), _np(n.weight), _np(n.reorder_index), n.variance_epsilon)
    hdq = _check_quant_stage(h, _np(q.post_attention_layernorm(x.float())), "rmsnorm+quant")
    gate, up = _gemm(h, real.mlp.gate_proj), _gemm(h, real.mlp.up_proj)
    _check_gemm_stage(gate, hdq, real.mlp.gate_proj, q.mlp.gate_proj, "gate_proj")
    _check_gemm_stage(up, hdq, real.mlp.up_proj, q.mlp.up_proj, "up_proj")
    act = O.activate_fp16_i4(gate, up)
    g32, u32 = torch.from_numpy(gate.astype(np.float32)), torch.from_numpy(up.astype(np.float32))
    adq = _check_quant_stage(act, _np(q.mlp.act_quant(q.mlp.act_fn(g32) * u32)), "silu*up+quant")
    out = _gemm(act, real.mlp.down_proj)
    _check_gemm_stage(out, adq, real.mlp.down_proj, q.mlp.down_proj, "down_proj")
    # end to end: simulator alone vs kernels alone
    sim = _np(q.mlp(q.post_attention_layernorm(x.float()[None]))[0])
    e2e = np.abs(out.astype(np.float32) - sim).max() / np.abs(sim).max()
    assert e2e < 0.05, e2e
    # the folded down_proj order is load-bearing: undo gate's row permutation and the result is garbage
    wrong = _gemm(O.activate_fp16_i4(np.ascontiguousarray(gate[:, ::-1]), up), real.mlp.down_proj).astype(np.float32)
    assert np.abs(wrong - sim).max() / np.abs(sim).max() > 10 * e2e


def test_exported_attention_projections_stage_by_stage():
    q, a = _build(seed=1)
    real = int4_decoder_layer(q, device=None)
    x = _input(6, 256, 2)
    n = real.input_layernorm
    h = O.rmsnorm_fp16_i4(_np(x), _np(n.weight), _np(n.reorder_index), n.variance_epsilon)
    hdq_np = _check_quant_stage(h, _np(q.input_layernorm(x.float())), "rmsnorm+quant")
    hdq = torch.from_numpy(hdq_np)
    _check_gemm_stage(_gemm(h, real.self_attn.q_proj), hdq_np, real.self_attn.q_proj, q.self_attn.q_proj, "q_proj")
    # K / V: the INT4 epilogue stores nibble*scale - zero per (token, head).  With min/max over v (signed_minmax, the
    # mathematically intended variant) every element is within half a step of the FP product; the reference kernel --
    # which the CUDA path reproduces bit for bit -- takes min/max over |v| (o4 epilogue), so only v >= min|v| is
    # representable and negative values wrap around in the nibble: there the bound is asserted on the covered range.
    for name in ("k_proj", "v_proj"):
        lin = getattr(real.self_attn, name)
        wd, _ = _dequant_weight(lin)
        ref = (hdq_np.astype(np.float64) @ wd.astype(np.float64).T).reshape(6, 2, 128)
        slack = 1e-2 * np.abs(ref).max()
        for signed in (True, False):
            d, p = _gemm(h, lin, signed_minmax=signed)
            nib = np.stack((d & 0xF, d >> 4), -1).reshape(6, 2, 128).astype(np.float64)
            prm = p.astype(np.float64).reshape(6, 2, 2)
            deq = nib * prm[..., :1] - prm[..., 1:]
            covered = np.ones_like(ref, bool) if signed else (ref >= -prm[..., 1:] + slack)
            assert covered.mean() > 0.3
            assert (np.abs(deq - ref)[covered] <= 0.51 * np.broadcast_to(prm[..., :1], ref.shape)[covered] + slack).all(), (name, signed)
    # o_proj input path: reorder_fp16_i4 with self_attn.reorder_index == index_select + act_quant of the simulator
    attn = _input(6, 256, 3)
    r = O.reorder_fp16_i4(_np(attn), _np(real.self_attn.reorder_index))
    rdq = _check_quant_stage(r, _np(q.self_attn.act_quant(torch.index_select(attn.float(), 1, q.self_attn.reorder_index))), "reorder+quant")
    _check_gemm_stage(_gemm(r, real.self_attn.o_proj), rdq, real.self_attn.o_proj, q.self_attn.o_proj, "o_proj")


def test_int4_checkpoint_round_trip(tmp_path):
    """export -> save -> load: every operand identical, config and metadata preserved, wrong files refused."""
    from atom_b200.checkpoint import load_int4, save_int4
    from atom_b200.llama import LlamaDecoderLayer
    q, a = _build(seed=3)
    real = int4_decoder_layer(q, device=None, layer_idx=2)
    path = str(tmp_path / "layer.safetensors")
    save_int4(real, path, extra={"source": "toy", "w_clip_ratio": a.w_clip_ratio})
    back, extra = load_int4(path, device="cpu")
    assert isinstance(back, LlamaDecoderLayer) and back.self_attn.layer_idx == 2 and extra["source"] == "toy"
    sd0, sd1 = real.state_dict(), back.state_dict()
    assert sd0.keys() == sd1.keys() and len(sd0) == 7 * 4 + 2 * 2 + 1
    for k in sd0:
        assert sd0[k].dtype == sd1[k].dtype and torch.equal(sd0[k], sd1[k]), k
    assert not any(p.is_meta for p in back.parameters())
    assert back.input_layernorm.variance_epsilon == real.input_layernorm.variance_epsilon
    # a foreign safetensors file is refused by its format tag
    from safetensors.torch import save_file
    other = str(tmp_path / "other.safetensors")
    save_file({"x": torch.zeros(2)}, other)
    import pytest
    with pytest.raises(ValueError, match="not an atom_b200"):
        load_int4(other, device="cpu")


def test_int4_checkpoint_full_model(tmp_path):
    from atom_b200.checkpoint import load_int4, save_int4
    from atom_b200.llama import LinearInt4, LlamaConfig, LlamaForCausalLM
    m = LlamaForCausalLM(LlamaConfig(hidden_size=256, intermediate_size=512, num_attention_heads=2, num_hidden_layers=2, vocab_size=64))
    for i, lin in enumerate(x for x in m.modules() if isinstance(x, LinearInt4)):
        lin.init_random(i)
    path = str(tmp_path / "model.safetensors")
    save_int4(m, path)
    back, _ = load_int4(path, device="cpu")
    assert back.model.config.num_hidden_layers == 2 and back.model.layers[1].self_attn.layer_idx == 1
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), back.state_dict().values()))
