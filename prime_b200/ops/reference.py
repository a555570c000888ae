"""Plain-PyTorch reference implementations of every hot op.

These are (a) the CPU/gloo execution path (BASELINE.json config 1) and (b) the
fp32 numerical oracle every sm_100a kernel is tested against
(``tests/test_kernels_gpu.py``).  Nothing here is ever selected on a CUDA
tensor by :mod:`prime_b200.ops.functional` — on the GPU the native kernels run
or the call fails loudly.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    xf = x.float()
    rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (xf * rstd * weight.float()).to(x.dtype)


def add_rmsnorm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float = 1e-5):
    """h = x + residual ; y = rmsnorm(h).  Returns (y, h)."""
    h = (x.float() + residual.float()).to(x.dtype)
    return rmsnorm(h, weight, eps), h


def linear(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    return F.linear(x, weight)


def rope_tables(seq_len: int, head_dim: int, theta: float = 10000.0, device=None) -> tuple[torch.Tensor, torch.Tensor]:
    """cos/sin tables [seq, head_dim/2] (fp32), interleaved-pair convention (Llama reference)."""
    freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 2, device=device, dtype=torch.float32) / head_dim))
    t = torch.arange(seq_len, device=device, dtype=torch.float32)
    ang = torch.outer(t, freqs)
    return ang.cos(), ang.sin()


def rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x: [B, S, H, D]; rotates interleaved pairs (x[2i], x[2i+1])."""
    B, S, H, D = x.shape
    xf = x.float().reshape(B, S, H, D // 2, 2)
    c = cos[:S].view(1, S, 1, D // 2)
    s = sin[:S].view(1, S, 1, D // 2)
    x0, x1 = xf[..., 0], xf[..., 1]
    out = torch.stack((x0 * c - x1 * s, x0 * s + x1 * c), dim=-1)
    return out.reshape(B, S, H, D).to(x.dtype)


def swiglu(gate_up: torch.Tensor) -> torch.Tensor:
    """gate_up: [..., 2*F] = concat(gate, up) → silu(gate) * up."""
    g, u = gate_up.float().chunk(2, dim=-1)
    return (F.silu(g) * u).to(gate_up.dtype)


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = True) -> torch.Tensor:
    """q: [B, S, H, D]; k/v: [B, S, Hkv, D] → [B, S, H, D]."""
    B, S, H, D = q.shape
    Hkv = k.shape[2]
    qf, kf, vf = (t.float().transpose(1, 2) for t in (q, k, v))
    if Hkv != H:
        kf = kf.repeat_interleave(H // Hkv, dim=1)
        vf = vf.repeat_interleave(H // Hkv, dim=1)
    scores = qf @ kf.transpose(-1, -2) / math.sqrt(D)
    if causal:
        mask = torch.ones(S, S, dtype=torch.bool, device=q.device).tril()
        scores = scores.masked_fill(~mask, float("-inf"))
    out = scores.softmax(-1) @ vf
    return out.transpose(1, 2).to(q.dtype)


def cross_entropy(logits: torch.Tensor, targets: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    return F.cross_entropy(logits.float().view(-1, logits.shape[-1]), targets.view(-1), ignore_index=ignore_index)


def adamw_step(
    p: torch.Tensor,
    g: torch.Tensor,
    m: torch.Tensor,
    v: torch.Tensor,
    *,
    lr: float,
    beta1: float,
    beta2: float,
    eps: float,
    weight_decay: float,
    step: int,
    grad_scale: float = 1.0,
) -> None:
    """In-place decoupled AdamW on fp32 master tensors (torch.optim.AdamW semantics)."""
    g = g.float() * grad_scale
    p.mul_(1.0 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    bc1 = 1.0 - beta1**step
    bc2 = 1.0 - beta2**step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def quantize_int8_blockwise(x: torch.Tensor, block: int = 1024) -> tuple[torch.Tensor, torch.Tensor]:
    """Symmetric per-block int8: returns (q int8 [n], scales fp32 [ceil(n/block)])."""
    n = x.numel()
    nb = (n + block - 1) // block
    pad = nb * block - n
    xf = x.float().reshape(-1)
    if pad:
        xf = F.pad(xf, (0, pad))
    xb = xf.view(nb, block)
    amax = xb.abs().amax(dim=1)
    # tensor ÷ tensor: ``amax / 127.0`` (tensor ÷ Python scalar) is evaluated by torch as amax · (1/127), which differs from the
    # IEEE quotient the kernel computes by one ulp for some blocks — enough to round a tie the other way in a few elements per
    # million and make the fused-vs-reference comparison data dependent
    scale = torch.div(amax, torch.full_like(amax, 127.0))
    inv = torch.where(scale > 0, torch.ones_like(scale) / scale, torch.zeros_like(scale))
    q = torch.clamp(torch.round(xb * inv[:, None]), -127, 127).to(torch.int8)
    return q.view(-1)[:n].contiguous(), scale


def dequantize_int8_blockwise(q: torch.Tensor, scale: torch.Tensor, block: int = 1024) -> torch.Tensor:
    n = q.numel()
    nb = scale.numel()
    pad = nb * block - n
    qf = q.float()
    if pad:
        qf = F.pad(qf, (0, pad))
    return (qf.view(nb, block) * scale[:, None]).view(-1)[:n]


def nesterov_outer_step(
    theta0: torch.Tensor,
    avg_pseudo_grad: torch.Tensor,
    momentum_buf: torch.Tensor,
    *,
    lr: float,
    momentum: float,
    nesterov: bool = True,
) -> None:
    """torch.optim.SGD(nesterov=True) semantics, in place on fp32 theta0 / momentum."""
    momentum_buf.mul_(momentum).add_(avg_pseudo_grad)
    step = avg_pseudo_grad + momentum * momentum_buf if nesterov else momentum_buf
    theta0.add_(step, alpha=-lr)


# ----------------------------------------------------------------------------------------------------------- MXFP8
def _mxfp8_sf_index(rows: int, k: int, device=None) -> torch.Tensor:
    """Byte offset of the scale of (row r, 32-block kb) in the tensor-core layout used by csrc/gemm_mxfp8_sm100.cu:
    [k/128][ceil(rows/128)+1] atoms of 512 bytes; inside an atom (r % 32) * 16 + ((r % 128) // 32) * 4 + kb % 4."""
    atoms = (rows + 127) // 128 + 1
    r = torch.arange(rows, device=device).view(-1, 1)
    kb = torch.arange(k // 32, device=device).view(1, -1)
    return ((kb >> 2) * atoms + (r >> 7)) * 512 + (r & 31) * 16 + ((r & 127) >> 5) * 4 + (kb & 3)


def quantize_mxfp8(x: torch.Tensor, transpose: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
    """OCP MX e4m3: per 32 contiguous contraction elements a power-of-two scale 2^e, the smallest with amax / 2^e <= 448."""
    xf = (x.t() if transpose else x).float().contiguous()
    rows, k = xf.shape
    assert k % 128 == 0
    blocks = xf.view(rows, k // 32, 32)
    amax = blocks.abs().amax(dim=-1)
    v = (amax * (1.0 / 448.0)).to(torch.float32)
    bits = v.view(torch.int32)
    e = (bits >> 23) & 0xFF
    e = torch.where((bits & 0x7FFFFF) != 0, e + 1, e).clamp(1, 254)
    inv = ((254 - e) << 23).to(torch.int32).view(torch.float32)
    q = (blocks * inv.unsqueeze(-1)).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8).view(rows, k)
    sf = torch.full(((k // 128) * ((rows + 127) // 128 + 1) * 512,), 127, dtype=torch.uint8, device=x.device)
    sf[_mxfp8_sf_index(rows, k, x.device).reshape(-1)] = e.to(torch.uint8).reshape(-1)
    return q, sf


def dequantize_mxfp8(q: torch.Tensor, sf: torch.Tensor) -> torch.Tensor:
    rows, k = q.shape
    e = sf[_mxfp8_sf_index(rows, k, q.device).reshape(-1)].view(rows, k // 32).to(torch.int32)
    scale = (e << 23).view(torch.float32)
    return (q.view(torch.float8_e4m3fn).float().view(rows, k // 32, 32) * scale.unsqueeze(-1)).view(rows, k)
