"""CPU suite: the model/ operator surface (atom_b200.quant / qlinear / qllama / qmixtral) against fixtures produced by
the reference's own Python (tests/golden/ref_py_fakequant.npz <- model/quant.py + qLinearLayer.py) and against float64
restatements.  These paths are torch code (offline weight preparation and the accuracy simulator); the real-INT4 forward
of QLinearLayer.pack() needs a GPU and is covered in test_gpu_layers.py."""
import os
import types

import numpy as np
import pytest
import torch

from atom_b200 import quant as Q
from atom_b200.qlinear import QLinearLayer, find_qlinear_layers


def _args(**kw):
    d = dict(wbits=4, abits=4, w_sym=True, a_sym=True, weight_group_size=128, act_group_size=128, weight_channel_group=2,
             w_clip_ratio=0.85, a_clip_ratio=0.9, keeper=128, keeper_precision=3, exponential=False, tiling=0, quant_type="int",
             static=False, kv_clip_ratio=1.0, reorder=True, kv_cache=True)
    d.update(kw)
    return types.SimpleNamespace(**d)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_py_fakequant.npz"))


def test_quant_functions_bit_match_reference_python(gold):
    a = _args()
    t = torch.from_numpy(gold["t"])
    assert torch.equal(Q.quantize_tensor(t.clone(), 4, 128, 0, True, 0.9), torch.from_numpy(gold["t_sym"]))
    assert torch.equal(Q.quantize_tensor(t.clone(), 4, 128, 0, False, 1.0), torch.from_numpy(gold["t_asym"]))
    assert torch.equal(Q.quantize_activation_wrapper(torch.from_numpy(gold["x"]).clone(), a), torch.from_numpy(gold["xq"]))
    assert torch.equal(Q.quantize_attn_k_wrapper(torch.from_numpy(gold["kv"]).clone(), a), torch.from_numpy(gold["kq"]))
    assert torch.equal(Q.quantize_attn_v_wrapper(torch.from_numpy(gold["kv"]).clone(), a), torch.from_numpy(gold["kq"]))


def test_qlinear_quant_matches_reference_and_pack_dequantises_to_it(gold):
    from oracle import oracle as O
    lin = torch.nn.Linear(512, 256, bias=False)
    lin.weight.data = torch.from_numpy(gold["w0"]).clone()
    plain = QLinearLayer(torch.nn.Linear(512, 256, bias=False), _args())
    plain.quant()
    with pytest.raises(RuntimeError, match="keep_fp_for_export"):                    # the FP weight is not kept by default
        plain.pack(device="cpu")
    q = QLinearLayer(lin, _args(keep_fp_for_export=True))
    q.quant()
    assert torch.equal(q.weight, torch.from_numpy(gold["wq"]))                       # fake-quant weight: bit exact
    y = q(Q.quantize_activation_wrapper(torch.from_numpy(gold["x"]).clone(), _args()))
    assert torch.allclose(y, torch.from_numpy(gold["y"]), rtol=1e-5, atol=1e-5)
    q.pack(device="cpu")
    assert q.weight_int4.shape == (256, 192) and q.weight_int4.dtype == torch.uint8
    assert q.weight_int8.shape == (256, 128) and q.scale_int4.shape == (3, 256) and q.scale_int8.shape == (256,)
    s4 = q.scale_int4.float().numpy()
    assert np.array_equal(s4[:, 0::2], s4[:, 1::2])                                   # weight_channel_group = 2
    w4 = O.unpack_int4(q.weight_int4.numpy()).astype(np.float32).reshape(256, 3, 128) * s4.T.reshape(256, 3, 1)
    w8 = q.weight_int8.numpy().astype(np.float32) * q.scale_int8.float().numpy()[:, None]
    wd = np.concatenate([w4.reshape(256, 384), w8], 1)
    assert np.abs(wd[:, :384] - gold["wq"][:, :384]).max() <= 1e-3 * np.abs(gold["wq"]).max()   # body: fp16 scale rounding only
    s8 = q.scale_int8.float().numpy()
    assert np.array_equal(s8[0::2], s8[1::2])                                        # keeper scale shared by channel pairs
    assert (np.abs(wd[:, 384:] - gold["wq"][:, 384:]) <= 1.05 * s8[:, None]).all()   # keeper: within one INT8 step of the per-row one
    assert list(find_qlinear_layers(torch.nn.Sequential(q))) == ["0"]
    with pytest.raises(RuntimeError):
        q(torch.randn(2, 512))        # packed layer runs the CUDA kernels only: CPU input is refused, no silent fallback


def test_qlinear_reorder_and_bf_passthrough():
    a = _args()
    lin = torch.nn.Linear(256, 8, bias=True)
    q = QLinearLayer(lin, a)
    idx = torch.randperm(256)
    w0 = q.weight.clone()
    q.reorder(idx)
    assert torch.equal(q.weight, w0[:, idx]) and q.bias is not None
    q16 = QLinearLayer(torch.nn.Linear(256, 8), _args(wbits=16))
    w = q16.weight.clone()
    q16.quant()
    assert torch.equal(q16.weight, w)


def test_quantizer_dynamic_and_static_paths():
    a = _args()
    qz = Q.Quantizer(a)
    x = torch.randn(3, 256)
    assert torch.equal(qz(x), x)                                  # unconfigured: identity
    qz.configure(Q.make_act_quant(a), None)
    assert torch.equal(qz(x.clone()), Q.quantize_activation_wrapper(x.clone(), a))
    s = _args(static=True, keeper=0)
    qs = Q.Quantizer(s)
    scales = torch.full((2, 1), 0.1)
    qs.configure(lambda t: t, scales)
    out = qs(torch.tensor([[0.26, -0.31] * 64 + [1.0, -2.0] * 64]).reshape(1, 256).clone())
    assert out.shape == (1, 256) and torch.allclose(out[0, :2], torch.tensor([0.3, -0.3]), atol=1e-6)
    assert out[0, 128].item() == pytest.approx(0.7)               # clamped to q_max * scale


def _toy_llama_layer(hidden=256, inter=512, heads=2):
    from atom_b200.qllama import ToyLlamaDecoderLayer
    torch.manual_seed(0)
    return ToyLlamaDecoderLayer(hidden, inter, heads)


def test_qllama_decoder_layer_surface_and_semantics():
    from atom_b200.qllama import QLlamaDecoderLayer
    a = _args(keeper=128)
    base = _toy_llama_layer()
    x = torch.randn(1, 5, 256)
    y_fp = base(x)[0]
    q = QLlamaDecoderLayer(base, _args(wbits=16, abits=16, kv_cache=False))
    assert torch.allclose(q(x)[0], y_fp, atol=1e-5)               # 16-bit: wrapping is the identity
    # attribute names that modelutils_llama.py:33-149 touches
    for name in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj", "self_attn.act_quant",
                 "self_attn.k_quant", "self_attn.v_quant", "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj", "mlp.act_quant",
                 "input_layernorm.originalNorm", "input_layernorm.act_quant", "post_attention_layernorm.act_quant"):
        obj = q
        for part in name.split("."):
            obj = getattr(obj, part)
    # W4A4: reorder with a permutation must not change the (un-quantised) function
    q4 = QLlamaDecoderLayer(_toy_llama_layer(), a)
    perm = torch.randperm(256)
    q4.self_attn.q_proj.reorder(perm); q4.self_attn.k_proj.reorder(perm); q4.self_attn.v_proj.reorder(perm)
    q4.input_layernorm.register_buffer("reorder_index", perm)
    assert torch.allclose(q4(x)[0], y_fp, atol=1e-4)
    # quantised: finite, close-ish to fp (4-bit noise), and quantisers are actually applied
    for m in find_qlinear_layers(q4).values():
        m.quant()
    f = Q.make_act_quant(a)
    q4.input_layernorm.act_quant.configure(f, None); q4.post_attention_layernorm.act_quant.configure(f, None)
    q4.self_attn.act_quant.configure(f, None); q4.mlp.act_quant.configure(f, None)
    q4.self_attn.k_quant.configure(lambda t: Q.quantize_attn_k_wrapper(t, a), None)
    q4.self_attn.v_quant.configure(lambda t: Q.quantize_attn_v_wrapper(t, a), None)
    y4 = q4(x)[0]
    assert torch.isfinite(y4).all() and not torch.allclose(y4, y_fp, atol=1e-4)
    assert (y4 - y_fp).abs().max() < 0.5 * y_fp.abs().max()


def test_qmixtral_layer_surface():
    from atom_b200.qmixtral import QMixtralDecoderLayer, ToyMixtralDecoderLayer
    torch.manual_seed(1)
    base = ToyMixtralDecoderLayer(hidden=256, inter=256, heads=2, kv_heads=1, experts=4, top_k=2)
    x = torch.randn(2, 3, 256)
    y_fp = base(x)[0]
    q = QMixtralDecoderLayer(base, _args(wbits=16, abits=16, kv_cache=False))
    assert torch.allclose(q(x)[0], y_fp, atol=1e-5)
    assert q.block_sparse_moe.gate.enable_quant is False          # router stays FP (qMixtralLayer.py:289)
    assert len(q.block_sparse_moe.experts) == 4 and hasattr(q.block_sparse_moe.experts[0], "act_quant")
    assert hasattr(q, "act_quant") and hasattr(q.self_attn, "k_quant")


def test_mixtral_drivers_reorder_is_function_preserving_and_router_stays_fp():
    from atom_b200 import modelutils
    from atom_b200.qmixtral import ToyMixtralDecoderLayer
    torch.manual_seed(2)
    a = _args(kv_cache=True)
    base = ToyMixtralDecoderLayer(hidden=256, inter=256, heads=2, kv_heads=1, experts=4, top_k=2)
    x = torch.randn(1, 6, 256)
    y_fp = base(x)[0]
    idx = {"layers.0.self_attn.k_proj.input": torch.randperm(256), "layers.0.self_attn.o_proj.input": torch.randperm(256),
           "layers.0.block_sparse_moe.experts.0.w1.input": torch.randperm(256),
           "layers.0.block_sparse_moe.experts.0.w2.input": torch.randperm(256)}
    layers = modelutils.reorder_model_mixtral([base], a, idx)
    q = layers[0]
    assert torch.allclose(q(x)[0], y_fp, atol=1e-4)              # a permutation applied consistently changes nothing
    gate_w = q.block_sparse_moe.gate.weight.clone()
    w1_fp = q.block_sparse_moe.experts[0].w1.weight.clone()
    modelutils.quantize_model_mixtral(layers, a)
    assert torch.equal(q.block_sparse_moe.gate.weight, gate_w)    # router untouched (modelutils_mixtral.py:139-145)
    assert not torch.equal(q.block_sparse_moe.experts[0].w1.weight, w1_fp)
    assert q.block_sparse_moe.experts[0].w1._w_unquantized is None     # no FP copy unless args.keep_fp_for_export
    modelutils.add_act_quant_wrapper_mixtral(layers, a)
    y4, router = q(x, output_router_logits=True)
    assert torch.isfinite(y4).all() and router.shape == (6, 4)
    # the router saw un-quantised activations: its logits equal gate(post_attention_layernorm(h)) of the FP hidden state
    assert (y4 - y_fp).abs().max() < 0.6 * y_fp.abs().max() and not torch.allclose(y4, y_fp, atol=1e-4)
