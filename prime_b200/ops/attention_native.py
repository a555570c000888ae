"""Flash-attention on sm_100a (tcgen05 QKᵀ / PV with TMEM accumulators).

``supported`` gates which shapes the native kernel covers; everything else is routed to the SDPA
fallback by :func:`prime_b200.ops.functional.attention` (outside the named hot path).
"""

from __future__ import annotations

import torch

from . import _lib


def supported(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> bool:
    try:
        lib = _lib.load()
    except Exception:
        return False
    if getattr(lib, "pb_flash_attn_fwd", None) is None:
        return False
    D = q.shape[-1]
    return q.dtype == torch.bfloat16 and D in (64, 128) and q.shape[1] % 128 == 0 and q.shape[2] % k.shape[2] == 0


def flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = True) -> torch.Tensor:
    raise NotImplementedError("native flash attention lands in a later milestone")
