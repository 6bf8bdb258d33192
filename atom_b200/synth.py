"""Synthetic random-quantised operands of the named shapes (BASELINE.json: "synthetic random-quantized tensors").

Pure torch, device-side where it matters; used by bench.py, __graft_entry__.smoke() and the model layers'
random initialisation (the reference's e2e harness also runs on random INT4 weights, e2e/README.md:9).
"""
import torch

from .ops import scale_size


def scale_index(row: torch.Tensor) -> torch.Tensor:
    """Reorder.cuh:39-44 on a tensor of row ids."""
    return (row // 16) * 64 + (row % 8) * 8 + (row // 8) % 2


def a_scale_layout(scales: torch.Tensor) -> torch.Tensor:
    """[G, M] per-(group,row) scales -> [G, scale_size(M)] fp16 in the ldmatrix-replicated layout (4 replicas)."""
    g, m = scales.shape
    out = torch.zeros((g, scale_size(m)), dtype=torch.float16, device=scales.device)
    base = scale_index(torch.arange(m, device=scales.device))
    for j in range(4):
        out[:, base + 2 * j] = scales.to(torch.float16)
    return out


def gemm_operands(m, n, k, device, seed=0):
    """a, b, a_scale, b_scale, a_keeper, b_keeper, a_keeper_scale, b_keeper_scale in the reference layouts
    (K includes the 128 keeper channels; B scales shared by adjacent output channels, as weight_channel_group=2 gives)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    g = k // 128 - 1
    kp = (k - 128) // 2
    a = torch.randint(0, 256, (m, kp), dtype=torch.uint8, device=device, generator=gen)
    b = torch.randint(0, 256, (n, kp), dtype=torch.uint8, device=device, generator=gen)
    ak = torch.randint(-128, 128, (m, 128), dtype=torch.int8, device=device, generator=gen)
    bk = torch.randint(-128, 128, (n, 128), dtype=torch.int8, device=device, generator=gen)
    sa = (torch.randn((g, m), device=device, generator=gen).abs() * 0.2 + 0.3) / 7.0
    sak = (torch.rand((1, m), device=device, generator=gen) * 16 + 8) / 127.0
    sb = 0.01 * (1 + torch.rand((g + 1, n // 2), device=device, generator=gen)).repeat_interleave(2, dim=1)
    return (a, b, a_scale_layout(sa), sb[:g].to(torch.float16).contiguous(), ak, bk, a_scale_layout(sak)[0].contiguous(),
            (sb[g] * 7.0 / 127.0).to(torch.float16).contiguous())
