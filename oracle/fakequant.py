"""CPU restatement of the reference's simulated-quantisation path (BASELINE config #1).

TEST INFRASTRUCTURE ONLY (checker + the `cpu_baseline` / `--impl reference` timing leg of bench.py).

Restates, in torch on the host cores:
  * quantize_tensor            /root/reference/model/quant.py:119-183  (uniform affine, sym & asym)
  * quantize_tensor_channel_group                        quant.py:69-107
  * quantize_activation_wrapper                          quant.py:188-231
  * QLinearLayer.quant / forward        /root/reference/model/qLinearLayer.py:32-77
Pinned against tests/golden/ref_py_fakequant.npz, which was produced by importing the reference's
own modules (tests/golden/make_golden.py).
"""
import types

import torch


def w4a4_args(**over):
    """The W4A4 recipe of README.md:89-95 / scripts/run_atom_ppl.sh as an args namespace."""
    d = dict(wbits=4, abits=4, w_sym=True, a_sym=True, weight_group_size=128, act_group_size=128,
             weight_channel_group=2, w_clip_ratio=0.85, a_clip_ratio=0.9, keeper=128, keeper_precision=3,
             exponential=False, tiling=0, quant_type="int", static=False, kv_clip_ratio=1.0, reorder=True)
    d.update(over)
    return types.SimpleNamespace(**d)


@torch.no_grad()
def fq_rows(w2d, n_bits, sym, clip_ratio=1.0):
    """Fake-quantise each row of a 2-D tensor with its own scale (quant.py:141-181, int branch)."""
    if sym:
        amax = w2d.abs().amax(dim=-1, keepdim=True).clamp(min=1e-5)
        qmax, qmin = 2 ** (n_bits - 1) - 1, -(2 ** (n_bits - 1))
        if clip_ratio < 1.0:
            amax = amax * clip_ratio
        scale = amax / qmax
        zero = torch.zeros_like(scale)
    else:
        hi, lo = w2d.amax(dim=-1, keepdim=True), w2d.amin(dim=-1, keepdim=True)
        qmax, qmin = 2 ** n_bits - 1, 0
        if clip_ratio < 1.0:
            hi, lo = hi * clip_ratio, lo * clip_ratio
        scale = (hi - lo).clamp(min=1e-5) / qmax
        zero = torch.round(-lo / scale).clamp_(min=qmin, max=qmax)
    return (torch.clamp(torch.round(w2d / scale) + zero, qmin, qmax) - zero) * scale


@torch.no_grad()
def fq_tensor(w, n_bits, group_size, sym, clip_ratio=1.0):
    shape = w.shape
    w2 = w.reshape(-1, group_size) if group_size > 0 else w.reshape(-1, shape[-1])
    return fq_rows(w2, n_bits, sym, clip_ratio).reshape(shape)


@torch.no_grad()
def fq_weight_channel_group(W, n_bits, group_size, sym, channel_group, clip_ratio):
    """Column groups of `group_size`; `channel_group` adjacent output rows share one scale (quant.py:80-105)."""
    W = W.clone()
    if group_size == 0:
        return fq_rows(W, n_bits, sym)
    for c0 in range(0, W.shape[1], group_size):
        blk = W[:, c0:c0 + group_size]
        if channel_group > 1:
            blk = blk.reshape(W.shape[0] // channel_group, -1)
        blk = fq_rows(blk.contiguous(), n_bits, sym, clip_ratio)
        W[:, c0:c0 + group_size] = blk.reshape(-1, group_size)
    return W


@torch.no_grad()
def fq_activation(x, args):
    """quant.py:188-231: INT8 per-row keeper on the last `keeper` channels, group INT4 elsewhere."""
    if args.abits >= 16:
        return x
    shape = x.shape
    x = x.reshape(-1, shape[-1]).clone()
    if args.keeper > 0:
        keep = x[:, -args.keeper:].clone()
        if args.keeper_precision == 3:
            keep = fq_rows(keep, 8, True)
        x[:, -args.keeper:] = 0
    x = fq_tensor(x, args.abits, args.act_group_size, args.a_sym, args.a_clip_ratio)
    if args.keeper > 0:
        x[:, -args.keeper:] = keep
    return x.reshape(shape)


@torch.no_grad()
def fq_linear_weight(weight, args):
    """qLinearLayer.py:42-77 (.quant())."""
    if args.wbits >= 16:
        return weight
    w = weight.clone()
    if args.keeper > 0:
        keep = w[:, -args.keeper:].clone().contiguous()
        if args.keeper_precision == 3:
            keep = fq_rows(keep, 8, True)
        w[:, -args.keeper:] = 0
    w = fq_weight_channel_group(w, args.wbits, args.weight_group_size, args.w_sym, args.weight_channel_group,
                                args.w_clip_ratio)
    if args.keeper > 0:
        w[:, -args.keeper:] = keep
    return w


@torch.no_grad()
def fq_linear_forward(x, wq, args):
    """One forward of the simulated W4A4 linear: act fake-quant + F.linear (qLlamaLayer.py:142-151 + qLinearLayer.py:33)."""
    return torch.nn.functional.linear(fq_activation(x, args), wq)
