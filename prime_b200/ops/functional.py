"""Public op surface.  CUDA tensors → hand-written sm_100a kernels (required, loud failure);
CPU tensors → the torch reference in :mod:`prime_b200.ops.reference`.

Every CUDA op is a ``torch.autograd.Function`` whose forward/backward launch kernels from the
native library on the *current* torch stream (so they compose with streams and CUDA graphs).
``launch_count()`` reports how many native kernels were launched — bench.py's ``gpu_launches``.
"""

from __future__ import annotations


import torch

from . import _lib, reference

_LAUNCHES = 0


def launch_count() -> int:
    return _LAUNCHES


def reset_launch_count() -> None:
    global _LAUNCHES
    _LAUNCHES = 0


def _count(n: int = 1) -> None:
    global _LAUNCHES
    _LAUNCHES += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


# --------------------------------------------------------------------------- GEMM


def gemm(
    a: torch.Tensor,
    b: torch.Tensor,
    *,
    a_mn_major: bool = False,
    b_mn_major: bool = False,
    out: torch.Tensor | None = None,
    out_dtype: torch.dtype = torch.bfloat16,
    accumulate: bool = False,
    max_ctas: int = 0,
) -> torch.Tensor:
    """C[M,N] (+)= A·Bᵀ on the tcgen05 kernel.

    ``a``: [M,K] (K-major) or, if ``a_mn_major``, stored [K,M].
    ``b``: [N,K] (K-major) or, if ``b_mn_major``, stored [K,N].
    Both must be 2-D bf16 with unit inner stride.
    """
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if a_mn_major:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn_major:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, f"contraction mismatch {K} vs {Kb}"
    if out is None:
        assert not accumulate
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    else:
        assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype in (torch.bfloat16, torch.float32)
    lib = _lib.load()
    rc = lib.pb_gemm_bf16(
        a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), b.stride(0), out.stride(0),
        int(a_mn_major), int(b_mn_major), int(out.dtype == torch.float32), int(accumulate), max_ctas, _stream(),
    )  # fmt: skip
    _lib.check(rc, "pb_gemm_bf16")
    _count()
    return out


def set_gemm_pair_mode(on: bool | None) -> int:
    """Tile scheduler of the tcgen05 GEMM: ``False`` = one CTA per 128×256 tile, ``True`` = a CTA pair (``cta_group::2``,
    the two SMs of a TPC) per 256×256 tile, ``None`` = re-read ``PB_GEMM_PAIR``. Returns the previous mode."""
    return int(_lib.load().pb_gemm_set_pair_mode(-1 if on is None else int(bool(on))))


def set_gemm_split_k(mode: int) -> int:
    """Split-K policy of fp32 reduce-add GEMMs (weight gradients): -1 automatic, 0 off, n>1 force n slices. Returns the old mode."""
    return int(_lib.load().pb_gemm_set_split_k(int(mode)))


def _take_fresh(param: torch.Tensor) -> bool:
    """True exactly once per optimizer step and parameter (the sharded engine arms ``_mg_fresh`` in ``zero_grad``): the first
    gradient write of the step OVERWRITES main_grad, which is what lets the engine skip the 4·N-byte memset."""
    if getattr(param, "_mg_fresh", False):
        param._mg_fresh = False
        return True
    return False


def _wgrad(dy2: torch.Tensor, x2: torch.Tensor, weight: torch.Tensor) -> torch.Tensor | None:
    """dW = dyᵀ·x: fused fp32 accumulation into ``weight.main_grad`` when the engine attached one, else a bf16 gradient."""
    main_grad = getattr(weight, "main_grad", None)
    if main_grad is not None:
        gemm(dy2, x2, a_mn_major=True, b_mn_major=True, out=main_grad, accumulate=not _take_fresh(weight))
        return None
    return gemm(dy2, x2, a_mn_major=True, b_mn_major=True)


# Which weight currently sits, complete, in each gather scratch region: full_ptr → (id of the RowShard, shard version).
# Everything runs on one stream in program order, so "the kernel that gathers it has been enqueued" is all that needs tracking.
_RESIDENT: dict[int, tuple[int, int]] = {}


def reset_gather_cache() -> None:
    """Forget what the gather scratch holds (CUDA-graph capture: a captured micro-step must not depend on what an earlier
    micro-step happened to leave resident; engine construction / teardown)."""
    _RESIDENT.clear()


def _resident(z) -> bool:
    return z.full_ptr != 0 and _RESIDENT.get(z.full_ptr) == (id(z), z.ver[0])


def _mark_resident(z) -> None:
    _RESIDENT[z.full_ptr] = (id(z), z.ver[0])


def _full_weight(z, device) -> torch.Tensor:
    if z.full is not None:
        return z.full
    return torch.as_tensor(_RawBuffer(z.full_ptr, z.rows * z.cols * 2), device=device).view(torch.bfloat16).view(z.rows, z.cols)


def gemm_wgather(a: torch.Tensor, z, *, b_mn_major: bool = False, out: torch.Tensor | None = None, rope: tuple | None = None,
                 swiglu_h: torch.Tensor | None = None, swiglu_bwd_gu: torch.Tensor | None = None) -> torch.Tensor:  # fmt: skip
    """Parameter all-gather ⊕ GEMM (ZeRO-3): ``z`` is the ``RowShard`` of a weight W [rows, cols] sharded by rows over the FSDP group.

    ``b_mn_major=False``: C[M, rows] = a[M, cols]·Wᵀ (forward; ``rope`` = RoPE epilogue, ``swiglu_h`` = SwiGLU epilogue with W = W13,
    C = gate_up, ``swiglu_h`` = h).  ``b_mn_major=True``: C[M, cols] = a[M, rows]·W (input gradient; with ``swiglu_bwd_gu`` = the
    saved gate_up the epilogue turns dh into d_gate_up [M, 2·cols] on the fly).

    The peers' row blocks are pulled over NVLink by copier warps INSIDE the GEMM kernel (csrc/gemm_sm100.cu, IO = 3). The same
    kernel also gathers AHEAD the weight the next GEMM of the pass will consume (``z.next_fwd`` / ``z.next_bwd``), so in steady
    state a kernel finds its own operand already resident (no gating, full speed) and only carries the next one's transfer."""
    import ctypes

    assert a.dtype == torch.bfloat16 and a.dim() == 2 and a.stride(1) == 1
    M = a.shape[0]
    N = z.cols if b_mn_major else z.rows
    assert a.shape[1] == (z.rows if b_mn_major else z.cols)
    if out is None:
        out = torch.empty((M, 2 * N if swiglu_bwd_gu is not None else N), dtype=torch.bfloat16, device=a.device)
    lib = _lib.load()
    epi = 3 if swiglu_bwd_gu is not None else (2 if swiglu_h is not None else (1 if rope is not None else 0))
    own_resident = _resident(z)
    nz = z.next_bwd if b_mn_major else z.next_fwd
    if nz is not None and (nz.full_ptr == 0 or _resident(nz) or nz.n != z.n):
        nz = None
    if M <= 128 or (own_resident and nz is None):
        # operand complete in the local scratch (gathered ahead by the previous kernel), or tiny M (unit tests, debug models: explicit
        # gather by peer loads first): the plain kernels
        if not own_resident:
            pp = _lib.PeerPtrs.of(z.peer_ptrs[i] for i in range(z.n))
            _lib.check(lib.pb_allgather_copy(ctypes.byref(pp), z.rpr * z.cols * 2, z.full_ptr, _stream()), "pb_allgather_copy")
            _count()
            _mark_resident(z)
        w = _full_weight(z, a.device)
        if epi == 3:
            if M > 128:
                rc = lib.pb_gemm_bf16_swiglu_bwd(a.data_ptr(), w.data_ptr(), swiglu_bwd_gu.data_ptr(), out.data_ptr(), M, z.cols, z.rows,
                                                 a.stride(0), w.stride(0), swiglu_bwd_gu.stride(0), out.stride(0), _stream())  # fmt: skip
                _lib.check(rc, "pb_gemm_bf16_swiglu_bwd")
                _count()
            else:
                dh = gemm(a, w, b_mn_major=True)
                _lib.check(lib.pb_swiglu_bwd(swiglu_bwd_gu.data_ptr(), dh.data_ptr(), out.data_ptr(), M, z.cols, _stream()), "pb_swiglu_bwd")
                _count()
        elif epi == 2:
            FF = z.rows // 2
            if M > 128 and FF % 64 == 0:
                rc = lib.pb_gemm_bf16_swiglu(a.data_ptr(), w.data_ptr(), out.data_ptr(), swiglu_h.data_ptr(), M, FF, z.cols, a.stride(0),
                                             w.stride(0), out.stride(0), swiglu_h.stride(0), _stream())  # fmt: skip
                _lib.check(rc, "pb_gemm_bf16_swiglu")
                _count()
            else:
                gemm(a, w, out=out)
                _lib.check(lib.pb_swiglu_fwd(out.data_ptr(), swiglu_h.data_ptr(), M, FF, _stream()), "pb_swiglu_fwd")
                _count()
        elif epi == 1:
            cos, sin, seq_len, rot_cols, head_dim = rope
            rc = lib.pb_gemm_bf16_rope(a.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, a.shape[1], a.stride(0), w.stride(0), out.stride(0), 0,
                                       cos.data_ptr(), sin.data_ptr(), seq_len, rot_cols, head_dim, _stream())  # fmt: skip
            _lib.check(rc, "pb_gemm_bf16_rope")
            _count()
        else:
            gemm(a, w, b_mn_major=b_mn_major, out=out)
        return out
    cos = sin = None
    seq_len, rot_cols, head_dim = 1, 0, 64
    if rope is not None:
        cos, sin, seq_len, rot_cols, head_dim = rope
        assert cos.dtype == torch.float32 and cos.is_contiguous() and sin.is_contiguous() and cos.shape[0] >= seq_len
    aux = swiglu_h if swiglu_h is not None else swiglu_bwd_gu
    rc = lib.pb_gemm_wgather(
        a.data_ptr(), z.peer_ptrs, z.n, z.rank, z.full_ptr, z.flags.data_ptr(), out.data_ptr(), _ptr(aux), M, z.rows, z.cols,
        a.stride(0), out.stride(0), aux.stride(0) if aux is not None else 0, int(b_mn_major), epi,
        _ptr(cos), _ptr(sin), seq_len, rot_cols, head_dim,
        int(own_resident), nz.peer_ptrs if nz is not None else None, nz.full_ptr if nz is not None else None,
        nz.rows if nz is not None else 0, nz.cols if nz is not None else 0, _stream(),
    )  # fmt: skip
    _lib.check(rc, "pb_gemm_wgather")
    _count(1 if own_resident else 2)
    _mark_resident(z)
    if nz is not None:
        _mark_resident(nz)
    return out


class _RawBuffer:
    """Raw device pointer → torch (zero copy) through ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


class _LinearFn(torch.autograd.Function):
    """y = x Wᵀ.  Weight gradient accumulates straight into ``weight.main_grad`` (fp32, fused in the GEMM
    epilogue) when the FSDP engine has attached one; otherwise a bf16 gradient is returned."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, rope: tuple | None = None) -> torch.Tensor:
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        ctx.save_for_backward(x2, weight)
        ctx.x_shape = x.shape
        # allocate with the final shape: a Function output must not be a view (RoPE / the loss rotate / overwrite it in place)
        y = torch.empty((*x.shape[:-1], weight.shape[0]), dtype=x.dtype, device=x.device)
        z = getattr(weight, "z3", None)
        if z is not None:  # ZeRO-3: the weight is gathered from the peers' shards inside the GEMM
            gemm_wgather(x2, z, out=y.view(-1, weight.shape[0]), rope=rope)
        elif rope is None:
            gemm(x2, weight, out=y.view(-1, weight.shape[0]))
        else:
            # fused QKV projection: the GEMM epilogue rotates the Q/K head columns while the tile is in registers. The consumer
            # (rope_attention_qkv(..., pre_rotated=True)) returns the gradient w.r.t. the UN-rotated activation, so backward below
            # is the plain linear backward.
            cos, sin, seq_len, rot_cols, head_dim = rope
            M, N, K = x2.shape[0], weight.shape[0], x2.shape[1]
            assert cos.dtype == torch.float32 and cos.is_contiguous() and sin.is_contiguous() and cos.shape[0] >= seq_len
            rc = _lib.load().pb_gemm_bf16_rope(x2.data_ptr(), weight.data_ptr(), y.data_ptr(), M, N, K, x2.stride(0), weight.stride(0), N, 0,
                                               cos.data_ptr(), sin.data_ptr(), seq_len, rot_cols, head_dim, _stream())  # fmt: skip
            _lib.check(rc, "pb_gemm_bf16_rope")
            _count()
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        x2, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(ctx.x_shape, dtype=dy.dtype, device=dy.device)
            z = getattr(weight, "z3", None)
            if z is not None:
                gemm_wgather(dy2, z, b_mn_major=True, out=dx.view(-1, ctx.x_shape[-1]))
            else:
                gemm(dy2, weight, b_mn_major=True, out=dx.view(-1, ctx.x_shape[-1]))
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dy2, x2, weight)
        return dx, dw, None


def linear(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    if x.is_cuda:
        return _LinearFn.apply(x, weight)
    return reference.linear(x, weight)


# A/B switch for measurements: PB_NO_EPILOGUE_FUSION=1 runs RoPE and SwiGLU as stand-alone kernels again
# --------------------------------------------------------------------------- block-scaled fp8 (MXFP8)


def mxfp8_sf_bytes(rows: int, k: int) -> int:
    """Size of the scale-factor buffer of a [rows, k] MXFP8 operand: [k/128][ceil(rows/128)+1] atoms of 512 bytes."""
    return (k // 128) * ((rows + 127) // 128 + 1) * 512


def quantize_mxfp8(x: torch.Tensor, transpose: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
    """bf16 [R, C] → (e4m3 values as uint8, UE8M0 scales in the tensor-core layout), one power-of-two scale per 32 elements.

    ``transpose=False``: values [R, C], blocks along C (operand contracted over C).
    ``transpose=True``:  values [C, R], blocks along R (the same matrix used with R as the contraction dimension)."""
    assert x.dim() == 2 and x.dtype == torch.bfloat16 and x.stride(1) == 1
    R, C = x.shape
    rows, k = (C, R) if transpose else (R, C)
    assert k % 128 == 0, f"MXFP8 contraction dimension must be a multiple of 128, got {k}"
    if not x.is_cuda:
        return reference.quantize_mxfp8(x, transpose)
    q = torch.empty((rows, k), dtype=torch.uint8, device=x.device)
    sf = torch.empty(mxfp8_sf_bytes(rows, k), dtype=torch.uint8, device=x.device)
    rc = _lib.load().pb_quantize_mxfp8(x.data_ptr(), x.stride(0), q.data_ptr(), sf.data_ptr(), R, C, int(transpose), _stream())
    _lib.check(rc, "pb_quantize_mxfp8")
    _count()
    return q, sf


def gemm_mxfp8(aq: torch.Tensor, asf: torch.Tensor, bq: torch.Tensor, bsf: torch.Tensor, out: torch.Tensor | None = None,
               max_ctas: int = 0) -> torch.Tensor:  # fmt: skip
    """C[M,N] bf16 = dequant(A)·dequant(B)ᵀ on ``tcgen05.mma.kind::mxf8f6f4.block_scale`` (scales applied by the tensor core)."""
    M, K = aq.shape
    N, Kb = bq.shape
    assert K == Kb and aq.dtype == torch.uint8 and bq.dtype == torch.uint8 and aq.is_contiguous() and bq.is_contiguous()
    if not aq.is_cuda:
        c = reference.dequantize_mxfp8(aq, asf) @ reference.dequantize_mxfp8(bq, bsf).t()
        if out is None:
            return c.to(torch.bfloat16)
        out.copy_(c)
        return out
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=aq.device)
    assert out.shape == (M, N) and out.dtype == torch.bfloat16 and out.stride(1) == 1
    rc = _lib.load().pb_gemm_mxfp8(aq.data_ptr(), asf.data_ptr(), bq.data_ptr(), bsf.data_ptr(), out.data_ptr(), M, N, K, out.stride(0),
                                   max_ctas, _stream())  # fmt: skip
    _lib.check(rc, "pb_gemm_mxfp8")
    _count()
    return out


def set_mxfp8_pair_mode(on: bool | None) -> int:
    """Scheduler of the MXFP8 GEMM: ``False`` = one CTA per 128×192 tile, ``True`` = CTA pair (``cta_group::2``) per 256×192 tile,
    ``None`` = re-read ``PB_MXFP8_PAIR``. Returns the previous mode."""
    return int(_lib.load().pb_gemm_mxfp8_set_pair_mode(-1 if on is None else int(bool(on))))


class _LinearMXFP8Fn(torch.autograd.Function):
    """y = x Wᵀ with the forward and the input-gradient GEMMs in block-scaled fp8; the weight gradient stays bf16 with fp32
    accumulation into ``main_grad`` (its contraction runs over tokens, where per-32 blocks along the token axis would need a
    third quantisation of both operands for the least precision-tolerant of the three GEMMs)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        ctx.save_for_backward(x2, weight)
        ctx.x_shape = x.shape
        xq, xsf = quantize_mxfp8(x2)
        wq, wsf = quantize_mxfp8(weight)
        y = torch.empty((*x.shape[:-1], weight.shape[0]), dtype=x.dtype, device=x.device)
        gemm_mxfp8(xq, xsf, wq, wsf, out=y.view(-1, weight.shape[0]))
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        x2, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dyq, dysf = quantize_mxfp8(dy2)                     # [M, N], blocks along N
            wtq, wtsf = quantize_mxfp8(weight, transpose=True)  # [K, N], blocks along N
            dx = torch.empty(ctx.x_shape, dtype=dy.dtype, device=dy.device)
            gemm_mxfp8(dyq, dysf, wtq, wtsf, out=dx.view(-1, ctx.x_shape[-1]))
        if ctx.needs_input_grad[1]:
            if x2.is_cuda:
                dw = _wgrad(dy2, x2, weight)
            else:
                dw = (dy2.float().t() @ x2.float()).to(weight.dtype)
        return dx, dw


def mxfp8_usable(x: torch.Tensor, weight: torch.Tensor) -> bool:
    """Both contraction dimensions (in for the forward, out for the input gradient) must be multiples of 128."""
    return weight.shape[0] % 128 == 0 and weight.shape[1] % 128 == 0 and x.dtype == torch.bfloat16


def linear_mxfp8(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """``linear`` with MXFP8 forward / input-gradient GEMMs (``train.precision = "mxfp8"``); falls back to bf16 for shapes the
    block format cannot express."""
    if not mxfp8_usable(x, weight):
        return linear(x, weight)
    return _LinearMXFP8Fn.apply(x, weight)


_NO_EPILOGUE_FUSION = bool(int(__import__("os").environ.get("PB_NO_EPILOGUE_FUSION", "0")))
# SwiGLU backward inside the down-projection's dgrad epilogue (EPI = 3) vs dgrad + stand-alone kernel. Default = what measured faster
# in the step (DESIGN.md §1.6 has the numbers of both); PB_SWIGLU_BWD_FUSED=0/1 overrides.
_SWIGLU_BWD_FUSED = bool(int(__import__("os").environ.get("PB_SWIGLU_BWD_FUSED", "1")))


def rope_fusable(x: torch.Tensor, n_heads: int, n_kv_heads: int, head_dim: int) -> bool:
    """Can the QKV projection rotate in its epilogue AND the native attention consume/undo it? (CUDA, [B,S,·] input, S%128==0)"""
    if _NO_EPILOGUE_FUSION or not (x.is_cuda and x.dim() == 3 and x.dtype == torch.bfloat16):
        return False
    from . import attention_native as native

    return native.supported_shape(x.shape[1], head_dim, n_heads, n_kv_heads)


def linear_qkv_rope(x: torch.Tensor, weight: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, n_heads: int, n_kv_heads: int) -> torch.Tensor:
    """Fused QKV projection ⊕ RoPE: [B,S,dim] → [B,S,(H+2Hkv)·D] with the Q and K heads already rotated (GEMM epilogue).
    Must be consumed by ``rope_attention_qkv(..., pre_rotated=True)``, whose backward undoes the rotation in its own epilogues."""
    head_dim = weight.shape[0] // (n_heads + 2 * n_kv_heads)
    return _LinearFn.apply(x, weight, (cos, sin, x.shape[1], (n_heads + n_kv_heads) * head_dim, head_dim))


# --------------------------------------------------------------------------- embedding


class _EmbeddingFn(torch.autograd.Function):
    """Row gather out of the parameter (ZeRO-3: out of the peers' shards over NVLink) and a deterministic, sort-based scatter-add of
    the gradient into the fp32 ``main_grad`` (csrc/embedding.cu) — no torch indexing / radix-sort / atomics kernels."""

    @staticmethod
    def forward(ctx, tokens: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        import ctypes

        lib = _lib.load()
        V, dim = weight.shape
        t = tokens.reshape(-1)
        t = t if t.is_contiguous() else t.contiguous()
        assert t.dtype == torch.int64
        z = getattr(weight, "z3", None)
        if z is not None:
            pp, rpr = _lib.PeerPtrs.of(z.peer_ptrs[i] for i in range(z.n)), z.rpr
        else:
            assert weight.is_contiguous()
            pp, rpr = _lib.PeerPtrs.of([weight.data_ptr()]), V
        out = torch.empty((*tokens.shape, dim), dtype=weight.dtype, device=weight.device)
        _lib.check(lib.pb_embedding_fwd(t.data_ptr(), t.numel(), ctypes.byref(pp), rpr, dim, out.data_ptr(), _stream()), "pb_embedding_fwd")
        _count()
        ctx.weight = weight
        ctx.sorted_ev = None
        if ctx.needs_input_grad[1]:  # (autograd runs forward() with grad mode off: needs_input_grad is the reliable signal)
            # the backward needs the positions sorted by token: a one-CTA kernel (~0.2 ms) that depends only on the token ids, so it
            # runs NOW on a side stream underneath the forward GEMMs instead of on the critical path at the end of the backward
            sorted_keys = torch.empty(t.numel(), dtype=torch.int64, device=weight.device)
            if torch.cuda.is_current_stream_capturing():
                _lib.check(lib.pb_embedding_sort(t.data_ptr(), t.numel(), sorted_keys.data_ptr(), _stream()), "pb_embedding_sort")
            else:
                side = _side_stream(weight.device)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    _lib.check(lib.pb_embedding_sort(t.data_ptr(), t.numel(), sorted_keys.data_ptr(), side.cuda_stream), "pb_embedding_sort")
                    ctx.sorted_ev = torch.cuda.Event()
                    ctx.sorted_ev.record(side)
                t.record_stream(side)
                sorted_keys.record_stream(side)
            _count((t.numel() + 16383) // 16384)
            ctx.save_for_backward(sorted_keys)
        return out

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        lib = _lib.load()
        if not ctx.saved_tensors:
            return None, None
        (sorted_keys,) = ctx.saved_tensors
        weight = ctx.weight
        V, dim = weight.shape
        d2 = dout.reshape(-1, dim)
        d2 = d2 if d2.is_contiguous() else d2.contiguous()
        main_grad = getattr(weight, "main_grad", None)
        if main_grad is not None:
            if _take_fresh(weight):
                main_grad.zero_()
            grad, ret = main_grad, None
        else:
            grad = torch.zeros((V, dim), dtype=torch.float32, device=dout.device)
            ret = grad
        if ctx.sorted_ev is not None:
            torch.cuda.current_stream().wait_event(ctx.sorted_ev)
        T = sorted_keys.numel()
        _lib.check(lib.pb_embedding_scatter(sorted_keys.data_ptr(), T, d2.data_ptr(), grad.data_ptr(), dim, _stream()), "pb_embedding_scatter")
        _count((T + 16383) // 16384)
        return None, (None if ret is None else ret.to(weight.dtype))


_SIDE_STREAMS: dict = {}


def _side_stream(device: torch.device) -> "torch.cuda.Stream":
    s = _SIDE_STREAMS.get(device)
    if s is None:
        s = _SIDE_STREAMS[device] = torch.cuda.Stream(device=device)
    return s


def embedding(tokens: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """``weight[tokens]`` for a bf16 [V, dim] table (dim % 8 == 0)."""
    if weight.is_cuda and weight.dtype == torch.bfloat16 and weight.shape[1] % 8 == 0:
        return _EmbeddingFn.apply(tokens, weight)
    return torch.nn.functional.embedding(tokens, weight)


# --------------------------------------------------------------------------- RMSNorm


def _accumulate_param_grad(param: torch.Tensor, grad_fp32: torch.Tensor | None):
    """Helper for ops that produce fp32 weight grads."""
    return None if grad_fp32 is None else grad_fp32.to(param.dtype)


class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, eps):
        lib = _lib.load()
        D = x.shape[-1]
        x2 = x.reshape(-1, D)
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        R = x2.shape[0]
        y = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        rstd = torch.empty(R, dtype=torch.float32, device=x.device)
        if residual is not None:
            r2 = residual.reshape(-1, D)
            r2 = r2 if r2.is_contiguous() else r2.contiguous()
            h = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        else:
            r2, h = None, x2
        rc = lib.pb_rmsnorm_fwd(_ptr(x2), _ptr(r2), _ptr(weight), _ptr(y), _ptr(h) if residual is not None else None,
                                _ptr(rstd), R, D, float(eps), _stream())  # fmt: skip
        _lib.check(rc, "pb_rmsnorm_fwd")
        _count()
        ctx.save_for_backward(h, weight, rstd)
        ctx.has_res = residual is not None
        ctx.shape = x.shape
        return y, (h if residual is not None else None)

    @staticmethod
    def backward(ctx, dy, dh):
        lib = _lib.load()
        h, weight, rstd = ctx.saved_tensors
        D = h.shape[-1]
        R = h.numel() // D
        dy2 = dy.reshape(R, D)
        dy2 = dy2 if dy2.is_contiguous() else dy2.contiguous()
        dres = None
        if dh is not None:
            dres = dh.reshape(R, D)
            dres = dres if dres.is_contiguous() else dres.contiguous()
        dx = torch.empty(ctx.shape, dtype=h.dtype, device=h.device)
        grid = lib.pb_rmsnorm_bwd_grid(R)
        partial = torch.empty((grid, D), dtype=torch.float32, device=h.device)
        main_grad = getattr(weight, "main_grad", None)
        if main_grad is not None:
            dw32, acc = main_grad, 0 if _take_fresh(weight) else 1
        else:
            dw32, acc = torch.empty(D, dtype=torch.float32, device=h.device), 0
        rc = lib.pb_rmsnorm_bwd(_ptr(dy2), _ptr(h), _ptr(weight), _ptr(rstd), _ptr(dres), _ptr(dx), _ptr(partial),
                                _ptr(dw32), acc, R, D, _stream())  # fmt: skip
        _lib.check(rc, "pb_rmsnorm_bwd")
        _count(2)
        dxv = dx
        dw = None if main_grad is not None else dw32.to(weight.dtype)
        # residual-add backward: gradient flows identically to x and residual
        return dxv, (dxv if ctx.has_res else None), dw, None


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    if x.is_cuda:
        return _RMSNormFn.apply(x, None, weight, eps)[0]
    return reference.rmsnorm(x, weight, eps)


def add_rmsnorm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float = 1e-5):
    """(rmsnorm(x + residual), x + residual) in one pass over HBM."""
    if x.is_cuda:
        return _RMSNormFn.apply(x, residual, weight, eps)
    return reference.add_rmsnorm(x, residual, weight, eps)


# --------------------------------------------------------------------------- RoPE (in place on fused QKV)


class _RopeQKVFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, cos, sin, n_heads, n_kv_heads, seq_len):
        # qkv: [B, S, (H + 2*Hkv) * D] contiguous, produced by the QKV GEMM → rotate in place
        lib = _lib.load()
        total_heads = n_heads + 2 * n_kv_heads
        D = qkv.shape[-1] // total_heads
        tokens = qkv.numel() // qkv.shape[-1]
        assert qkv.is_contiguous()
        rc = lib.pb_rope_inplace(_ptr(qkv), _ptr(cos), _ptr(sin), tokens, seq_len, n_heads + n_kv_heads, total_heads, D,
                                 1.0, _stream())  # fmt: skip
        _lib.check(rc, "pb_rope_inplace")
        _count()
        ctx.mark_dirty(qkv)
        ctx.save_for_backward(cos, sin)
        ctx.meta = (n_heads, n_kv_heads, seq_len, D, total_heads)
        return qkv

    @staticmethod
    def backward(ctx, dqkv):
        lib = _lib.load()
        cos, sin = ctx.saved_tensors
        n_heads, n_kv_heads, seq_len, D, total_heads = ctx.meta
        # never rotate the incoming gradient in place: autograd may share that tensor with other consumers
        dqkv = dqkv.clone(memory_format=torch.contiguous_format)
        tokens = dqkv.numel() // dqkv.shape[-1]
        # rotation is orthogonal: the backward is the inverse rotation, applied in place on the incoming grad
        rc = lib.pb_rope_inplace(_ptr(dqkv), _ptr(cos), _ptr(sin), tokens, seq_len, n_heads + n_kv_heads, total_heads, D,
                                 -1.0, _stream())  # fmt: skip
        _lib.check(rc, "pb_rope_inplace(bwd)")
        _count()
        return dqkv, None, None, None, None, None


def rope_qkv(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, n_heads: int, n_kv_heads: int) -> torch.Tensor:
    """Rotate the Q and K heads of a fused [B,S,(H+2Hkv)·D] activation."""
    B, S, W = qkv.shape
    if qkv.is_cuda:
        return _RopeQKVFn.apply(qkv, cos, sin, n_heads, n_kv_heads, S)
    D = W // (n_heads + 2 * n_kv_heads)
    x = qkv.view(B, S, n_heads + 2 * n_kv_heads, D)
    rot = reference.rope(x[:, :, : n_heads + n_kv_heads], cos, sin)
    return torch.cat((rot, x[:, :, n_heads + n_kv_heads :]), dim=2).view(B, S, W)


# --------------------------------------------------------------------------- SwiGLU


class _SwiGLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate_up):
        lib = _lib.load()
        F2 = gate_up.shape[-1]
        gu = gate_up.reshape(-1, F2)
        gu = gu if gu.is_contiguous() else gu.contiguous()
        out = torch.empty((*gate_up.shape[:-1], F2 // 2), dtype=gu.dtype, device=gu.device)
        rc = lib.pb_swiglu_fwd(_ptr(gu), _ptr(out), gu.shape[0], F2 // 2, _stream())
        _lib.check(rc, "pb_swiglu_fwd")
        _count()
        ctx.save_for_backward(gu)
        ctx.shape = gate_up.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        (gu,) = ctx.saved_tensors
        d2 = dout.reshape(gu.shape[0], -1)
        d2 = d2 if d2.is_contiguous() else d2.contiguous()
        dgu = torch.empty(ctx.shape, dtype=gu.dtype, device=gu.device)
        rc = lib.pb_swiglu_bwd(_ptr(gu), _ptr(d2), _ptr(dgu), gu.shape[0], gu.shape[1] // 2, _stream())
        _lib.check(rc, "pb_swiglu_bwd")
        _count()
        return dgu


def swiglu(gate_up: torch.Tensor) -> torch.Tensor:
    if gate_up.is_cuda:
        return _SwiGLUFn.apply(gate_up)
    return reference.swiglu(gate_up)


class _LinearSwiGLUFn(torch.autograd.Function):
    """h = silu(x W1ᵀ) · (x W3ᵀ) with W13 = [W1; W3]: ONE GEMM whose epilogue emits both the bf16 gate/up activations (kept for
    the backward) and h — the separate SwiGLU pass over the 370 MB gate_up tensor disappears from the forward. CTA-pair GEMM:
    the leader stages 128 gate rows of W13, its partner the matching 128 up rows."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, w13: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M, K = x2.shape
        FF = w13.shape[0] // 2
        gate_up = torch.empty((M, 2 * FF), dtype=x.dtype, device=x.device)
        h = torch.empty((*x.shape[:-1], FF), dtype=x.dtype, device=x.device)
        z = getattr(w13, "z3", None)
        if z is not None:
            gemm_wgather(x2, z, out=gate_up, swiglu_h=h.view(M, FF))
        else:
            rc = lib.pb_gemm_bf16_swiglu(x2.data_ptr(), w13.data_ptr(), gate_up.data_ptr(), h.data_ptr(), M, FF, K, x2.stride(0), w13.stride(0),
                                         2 * FF, FF, _stream())  # fmt: skip
            _lib.check(rc, "pb_gemm_bf16_swiglu")
            _count()
        ctx.save_for_backward(x2, w13, gate_up)
        ctx.x_shape = x.shape
        return h

    @staticmethod
    def backward(ctx, dh: torch.Tensor):
        lib = _lib.load()
        x2, w13, gate_up = ctx.saved_tensors
        M, FF = gate_up.shape[0], gate_up.shape[1] // 2
        d2 = dh.reshape(M, FF)
        d2 = d2 if d2.is_contiguous() else d2.contiguous()
        dgu = torch.empty_like(gate_up)
        rc = lib.pb_swiglu_bwd(_ptr(gate_up), _ptr(d2), _ptr(dgu), M, FF, _stream())
        _lib.check(rc, "pb_swiglu_bwd")
        _count()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(ctx.x_shape, dtype=dh.dtype, device=dh.device)
            z = getattr(w13, "z3", None)
            if z is not None:
                gemm_wgather(dgu, z, b_mn_major=True, out=dx.view(-1, ctx.x_shape[-1]))
            else:
                gemm(dgu, w13, b_mn_major=True, out=dx.view(-1, ctx.x_shape[-1]))
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dgu, x2, w13)
        return dx, dw


class _MLPSwiGLUFn(torch.autograd.Function):
    """y = (silu(x W1ᵀ)·(x W3ᵀ)) W2ᵀ as ONE autograd node so that the backward can fuse across the two projections:
    the down-projection's input-gradient GEMM (dh = dy·W2) applies the SwiGLU derivative in its epilogue and writes d_gate_up
    directly (``pb_gemm_bf16_swiglu_bwd``) — dh is never stored or re-read and the stand-alone swiglu_bwd pass is gone.
    Forward: gate/up GEMM with the SwiGLU epilogue, then the down projection. Saves exactly what the two separate nodes saved."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, w13: torch.Tensor, w2: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M, K = x2.shape
        FF = w13.shape[0] // 2
        gate_up = torch.empty((M, 2 * FF), dtype=x.dtype, device=x.device)
        h = torch.empty((M, FF), dtype=x.dtype, device=x.device)
        z13, z2 = getattr(w13, "z3", None), getattr(w2, "z3", None)
        if z13 is not None:
            gemm_wgather(x2, z13, out=gate_up, swiglu_h=h)
        else:
            rc = lib.pb_gemm_bf16_swiglu(x2.data_ptr(), w13.data_ptr(), gate_up.data_ptr(), h.data_ptr(), M, FF, K, x2.stride(0), w13.stride(0),
                                         2 * FF, FF, _stream())  # fmt: skip
            _lib.check(rc, "pb_gemm_bf16_swiglu")
            _count()
        y = torch.empty((*x.shape[:-1], w2.shape[0]), dtype=x.dtype, device=x.device)
        if z2 is not None:
            gemm_wgather(h, z2, out=y.view(M, -1))
        else:
            gemm(h, w2, out=y.view(M, -1))
        ctx.save_for_backward(x2, w13, w2, gate_up, h)
        ctx.x_shape = x.shape
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        lib = _lib.load()
        x2, w13, w2, gate_up, h = ctx.saved_tensors
        M, FF = h.shape
        dy2 = dy.reshape(M, -1)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dw2 = _wgrad(dy2, h, w2) if ctx.needs_input_grad[2] else None
        dgu = torch.empty_like(gate_up)
        z13, z2 = getattr(w13, "z3", None), getattr(w2, "z3", None)
        if not _SWIGLU_BWD_FUSED:
            # two kernels: dh = dy·W2 (plain / gather-fused dgrad), then the stand-alone SwiGLU backward (0.82 of the HBM roof)
            dh = torch.empty((M, FF), dtype=dy.dtype, device=dy.device)
            if z2 is not None:
                gemm_wgather(dy2, z2, b_mn_major=True, out=dh)
            else:
                gemm(dy2, w2, b_mn_major=True, out=dh)
            _lib.check(lib.pb_swiglu_bwd(_ptr(gate_up), _ptr(dh), _ptr(dgu), M, FF, _stream()), "pb_swiglu_bwd")
            _count()
        elif z2 is not None:
            gemm_wgather(dy2, z2, b_mn_major=True, out=dgu, swiglu_bwd_gu=gate_up)
        else:
            rc = lib.pb_gemm_bf16_swiglu_bwd(dy2.data_ptr(), w2.data_ptr(), gate_up.data_ptr(), dgu.data_ptr(), M, FF, dy2.shape[1],
                                             dy2.stride(0), w2.stride(0), gate_up.stride(0), dgu.stride(0), _stream())  # fmt: skip
            _lib.check(rc, "pb_gemm_bf16_swiglu_bwd")
            _count()
        dx = dw13 = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(ctx.x_shape, dtype=dy.dtype, device=dy.device)
            if z13 is not None:
                gemm_wgather(dgu, z13, b_mn_major=True, out=dx.view(M, -1))
            else:
                gemm(dgu, w13, b_mn_major=True, out=dx.view(M, -1))
        if ctx.needs_input_grad[1]:
            dw13 = _wgrad(dgu, x2, w13)
        return dx, dw13, dw2


def mlp_swiglu(x: torch.Tensor, w13: torch.Tensor, w2: torch.Tensor) -> torch.Tensor:
    """SwiGLU MLP ``linear(swiglu(linear(x, w13)), w2)``; on CUDA one autograd node with the SwiGLU forward AND backward fused into
    GEMM epilogues (M > 128, hidden % 64 == 0), otherwise the composition of the separate ops."""
    if (not _NO_EPILOGUE_FUSION and x.is_cuda and x.dtype == torch.bfloat16 and (w13.shape[0] // 2) % 64 == 0
            and x.numel() // x.shape[-1] > 128 and w2.stride(1) == 1 and w13.stride(1) == 1):  # fmt: skip
        return _MLPSwiGLUFn.apply(x, w13, w2)
    return linear(linear_swiglu(x, w13), w2)


def linear_swiglu(x: torch.Tensor, w13: torch.Tensor) -> torch.Tensor:
    """``swiglu(linear(x, w13))``; fused into one GEMM on CUDA when the shape allows it (M > 128, FF % 64 == 0)."""
    if not _NO_EPILOGUE_FUSION and x.is_cuda and x.dtype == torch.bfloat16 and (w13.shape[0] // 2) % 64 == 0 and x.numel() // x.shape[-1] > 128:
        return _LinearSwiGLUFn.apply(x, w13)
    return swiglu(linear(x, w13))


# --------------------------------------------------------------------------- cross entropy


class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, ignore_index, grad_scale, unit_upstream, loss_acc):
        lib = _lib.load()
        V = logits.shape[-1]
        z = logits.reshape(-1, V)
        assert z.is_contiguous()
        t = targets.reshape(-1)
        t = t if t.is_contiguous() else t.contiguous()
        R = z.shape[0]
        work = torch.empty(R + 4, dtype=torch.float32, device=z.device)  # [0:2] scales, [2] loss, [4:] per-row losses
        # valid-target count → gradient scale, the fused row kernel (logits are overwritten with d(mean loss)/d(logits): nothing else
        # needs them afterwards) and a fixed-order sum of the row losses — three native launches, no framework kernel
        rc = lib.pb_cross_entropy_loss(_ptr(z), _ptr(t), work[4:].data_ptr(), work.data_ptr(), work[2:].data_ptr(), _ptr(loss_acc), R, V,
                                       ignore_index, float(grad_scale), _stream())  # fmt: skip
        _lib.check(rc, "pb_cross_entropy_loss")
        _count(3)
        # NOTE: deliberately not mark_dirty(): the logits are consumed, not returned; their storage now holds the gradient
        ctx.save_for_backward(z)
        ctx.shape = logits.shape
        ctx.unit_upstream = unit_upstream
        return work[2].reshape(())

    @staticmethod
    def backward(ctx, dloss):
        (dz,) = ctx.saved_tensors
        # dz already holds the gradient for dloss == 1; rescale for loss scaling / grad accumulation
        if not ctx.unit_upstream:
            dz = dz * dloss.to(dz.dtype)
        return dz.view(ctx.shape), None, None, None, None, None


def cross_entropy(
    logits: torch.Tensor,
    targets: torch.Tensor,
    ignore_index: int = -100,
    *,
    grad_scale: float = 1.0,
    unit_upstream: bool = False,
    loss_acc: torch.Tensor | None = None,
) -> torch.Tensor:
    """Mean token cross-entropy.  On CUDA the logits buffer is consumed (overwritten with its gradient).

    ``grad_scale`` folds a gradient multiplier (e.g. 1/grad-accumulation-steps) into the fused kernel; the
    returned loss value is NOT scaled.  ``unit_upstream=True`` promises the loss is back-propagated with a
    unit upstream gradient, which removes one full pass over the logits gradient.  ``loss_acc``: optional fp32 scalar the
    (unscaled) loss is also added to, inside the finalising kernel — the trainer's per-step running sum.
    """
    if logits.is_cuda:
        return _CrossEntropyFn.apply(logits, targets, ignore_index, float(grad_scale), bool(unit_upstream), loss_acc)
    loss = reference.cross_entropy(logits, targets, ignore_index)
    if loss_acc is not None:
        loss_acc += loss.detach().float()
    if grad_scale != 1.0:
        # same contract on CPU: value unscaled, gradient scaled
        loss = loss.detach() + (loss - loss.detach()) * grad_scale
    return loss


# --------------------------------------------------------------------------- attention


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = True) -> torch.Tensor:
    """Unfused-layout attention, q [B,S,H,D], k/v [B,S,Hkv,D] → [B,S,H,D].  SDPA on CUDA (not a hot path: the models
    call :func:`attention_qkv`), torch reference on CPU."""
    if not q.is_cuda:
        return reference.attention(q, k, v, causal)
    import torch.nn.functional as F

    H, Hkv = q.shape[2], k.shape[2]
    out = F.scaled_dot_product_attention(
        q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=causal, enable_gqa=(H != Hkv)
    )
    return out.transpose(1, 2)


def attention_qkv(qkv: torch.Tensor, n_heads: int, n_kv_heads: int, causal: bool = True, impl: str = "auto") -> torch.Tensor:
    """Attention over the fused, RoPE-rotated QKV activation [B,S,(H+2Hkv)·D] → [B,S,H·D].

    ``impl``: "native" = tcgen05 flash-attention kernels (error if the shape is unsupported); "sdpa" = library
    fallback; "auto" = native whenever supported.
    """
    B, S, W = qkv.shape
    D = W // (n_heads + 2 * n_kv_heads)
    if qkv.is_cuda and impl in ("auto", "native"):
        from . import attention_native as native

        if native.supported(qkv, n_heads, n_kv_heads):
            return native.flash_attention_qkv(qkv, n_heads, n_kv_heads, causal)
        if impl == "native":
            raise RuntimeError(f"native flash attention does not support S={S}, D={D}, H={n_heads}, Hkv={n_kv_heads}")
    x = qkv.view(B, S, n_heads + 2 * n_kv_heads, D)
    q, k, v = x[:, :, :n_heads], x[:, :, n_heads : n_heads + n_kv_heads], x[:, :, n_heads + n_kv_heads :]
    return attention(q, k, v, causal).reshape(B, S, n_heads * D)


def rope_attention_qkv(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, n_heads: int, n_kv_heads: int, causal: bool = True,
                       impl: str = "auto", pre_rotated: bool = False) -> torch.Tensor:  # fmt: skip
    """RoPE on the Q/K heads of the fused QKV activation followed by attention → [B,S,H·D]. On the native path both run as one
    autograd node whose backward folds the inverse rotation into the dQ/dK epilogues. ``pre_rotated=True``: the producer
    (:func:`linear_qkv_rope`) already rotated in its GEMM epilogue; only the backward un-rotation is needed (native path only)."""
    if qkv.is_cuda and impl in ("auto", "native"):
        from . import attention_native as native

        if native.supported(qkv, n_heads, n_kv_heads):
            return native.rope_flash_attention_qkv(qkv, cos, sin, n_heads, n_kv_heads, causal, pre_rotated)
    if pre_rotated:
        raise RuntimeError("pre_rotated=True needs the native attention path (its backward undoes the rotation)")
    return attention_qkv(rope_qkv(qkv, cos, sin, n_heads, n_kv_heads), n_heads, n_kv_heads, causal, impl)
