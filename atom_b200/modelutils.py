"""Layer-list drivers of the model/ surface: /root/reference/model/modelutils_llama.py:14-149 (reorder_model_llama,
add_act_quant_wrapper_llama, quantize_model_llama) restated over a plain list of decoder layers, RTN weights only (the
GPTQ solver and the evaluation harness are outside the hot path)."""
from functools import partial

import torch

from .qlinear import find_qlinear_layers
from .qllama import QLlamaDecoderLayer
from .quant import quantize_activation_wrapper, quantize_attn_k_wrapper, quantize_attn_v_wrapper

_T4 = "layers.{}.{}.{}.{}"      # layers.10.self_attn.q_proj.input
_T3 = "layers.{}.{}.{}"


def _wrap(layer, args):
    return layer if isinstance(layer, QLlamaDecoderLayer) else QLlamaDecoderLayer(layer, args)


@torch.no_grad()
def reorder_model_llama(layers, args, reorder_index):
    """Permute every projection's input channels by its calibrated index; down_proj's order is folded into gate/up's
    *output* order; the norms (and the attention output) carry the index their consumers were permuted with."""
    assert reorder_index is not None, "Reorder index is None"
    for i in range(len(layers)):
        m = _wrap(layers[i], args)
        down = reorder_index[_T4.format(i, "mlp", "down_proj", "input")]
        m.mlp.gate_proj.reorder(reorder_index[_T4.format(i, "mlp", "gate_proj", "input")], down)
        m.mlp.up_proj.reorder(reorder_index[_T4.format(i, "mlp", "up_proj", "input")], down)
        m.mlp.down_proj.reorder(down, None)
        for p in ("q_proj", "k_proj", "v_proj", "o_proj"):     # outputs stay put: RoPE / head structure
            getattr(m.self_attn, p).reorder(reorder_index[_T4.format(i, "self_attn", p, "input")], None)
        m.input_layernorm.register_buffer("reorder_index", reorder_index[_T4.format(i, "self_attn", "k_proj", "input")])
        m.post_attention_layernorm.register_buffer("reorder_index", reorder_index[_T4.format(i, "mlp", "gate_proj", "input")])
        m.self_attn.register_buffer("reorder_index", reorder_index[_T4.format(i, "self_attn", "o_proj", "input")])
        layers[i] = m
    return layers


@torch.no_grad()
def add_act_quant_wrapper_llama(layers, args, scales=None):
    scales = scales or {}
    act = partial(quantize_activation_wrapper, args=args)
    for i in range(len(layers)):
        m = _wrap(layers[i], args)
        m.self_attn.act_quant.configure(act, scales.get(_T3.format(i, "self_attn", "o_proj")))
        m.self_attn.v_quant.configure(partial(quantize_attn_v_wrapper, args=args), None)
        m.self_attn.k_quant.configure(partial(quantize_attn_k_wrapper, args=args), None)
        m.mlp.act_quant.configure(act, scales.get(_T3.format(i, "mlp", "down_proj")))
        m.input_layernorm.act_quant.configure(act, scales.get(_T3.format(i, "self_attn", "k_proj")))
        m.post_attention_layernorm.act_quant.configure(act, scales.get(_T3.format(i, "mlp", "gate_proj")))
        layers[i] = m
    return layers


@torch.no_grad()
def quantize_model_llama(layers, args):
    for i in range(len(layers)):
        m = _wrap(layers[i], args)
        for q in find_qlinear_layers(m).values():
            q.quant()
        layers[i] = m
    return layers


# ------------------------------------------------------------------------------------------------ Mixtral
# /root/reference/model/modelutils_mixtral.py:14-150.  Differences from Llama the reference has, kept: q/k/v all take
# k_proj's input order; every expert shares expert 0's orders (w1/w3 input = experts.0.w1.input, w1/w3 output = w2 input =
# experts.0.w2.input) and so do the router and post_attention_layernorm; quantisers are assigned as plain callables; the
# router is never quantised.
_T6 = "layers.{}.{}.{}.{}.{}.{}"  # layers.10.block_sparse_moe.experts.0.w1.input


def _wrap_mixtral(layer, args):
    from .qmixtral import QMixtralDecoderLayer
    return layer if isinstance(layer, QMixtralDecoderLayer) else QMixtralDecoderLayer(layer, args)


@torch.no_grad()
def reorder_model_mixtral(layers, args, reorder_index):
    assert reorder_index is not None, "Reorder index is None"
    for i in range(len(layers)):
        m = _wrap_mixtral(layers[i], args)
        qkv = reorder_index[_T4.format(i, "self_attn", "k_proj", "input")]
        o_in = reorder_index[_T4.format(i, "self_attn", "o_proj", "input")]
        m.input_layernorm.register_buffer("reorder_index", qkv)
        for p in ("q_proj", "k_proj", "v_proj"):
            getattr(m.self_attn, p).reorder(qkv, None)
        m.self_attn.o_proj.reorder(o_in, None)
        m.self_attn.register_buffer("reorder_index", o_in)
        w1_in = reorder_index[_T6.format(i, "block_sparse_moe", "experts", 0, "w1", "input")]
        w2_in = reorder_index[_T6.format(i, "block_sparse_moe", "experts", 0, "w2", "input")]
        m.block_sparse_moe.gate.reorder(w1_in, None)
        for expert in m.block_sparse_moe.experts:
            expert.w1.reorder(w1_in, w2_in)
            expert.w3.reorder(w1_in, w2_in)
            expert.w2.reorder(w2_in, None)
        m.post_attention_layernorm.register_buffer("reorder_index", w1_in)
        layers[i] = m
    return layers


@torch.no_grad()
def add_act_quant_wrapper_mixtral(layers, args, scales=None):
    act = partial(quantize_activation_wrapper, args=args)
    for i in range(len(layers)):
        m = _wrap_mixtral(layers[i], args)
        m.self_attn.act_quant = act
        m.self_attn.v_quant = partial(quantize_attn_v_wrapper, args=args)
        m.self_attn.k_quant = partial(quantize_attn_k_wrapper, args=args)
        for expert in m.block_sparse_moe.experts:
            expert.act_quant = act
        m.act_quant = act
        m.block_sparse_moe.act_quant = act
        layers[i] = m
    return layers


@torch.no_grad()
def quantize_model_mixtral(layers, args):
    for i in range(len(layers)):
        m = _wrap_mixtral(layers[i], args)
        for expert in m.block_sparse_moe.experts:
            expert.quant()
        for p in ("q_proj", "k_proj", "v_proj", "o_proj"):
            getattr(m.self_attn, p).quant()
        layers[i] = m
    return layers
