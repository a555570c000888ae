"""Sharded inner-loop engine (the FSDP level of the two-level mesh).

Memory plan for 180 GB HBM3e parts: the bf16 parameters stay *resident and replicated* inside a
worker (one flat buffer), while everything optimizer-side — fp32 master weights, AdamW moments, the
reduced gradient — is partitioned 1/F per rank.  Per optimizer step the traffic is exactly FSDP's
(one gradient reduce-scatter + one parameter all-gather per bucket); what differs is *how*:

``backend="fused"`` (CUDA, symmetric NVLink heap — the product path)
    bucket ready in backward → flag signalled to peers → ``grad_reduce`` kernel pulls the peers'
    slices through NVSwitch (fixed rank order), scales, writes the fp32 shard gradient and the
    sum-of-squares partial;  after backward one ``adamw_push`` kernel per bucket does
    clip ⊕ AdamW ⊕ bf16 cast ⊕ *stores the new parameters straight into every peer's buffer*
    (the all-gather), then a flag barrier.  No NCCL call, no host sync, all on a side stream
    overlapped with the backward pass.

``backend="collective"`` (NCCL or gloo — the baseline "B0" and the CPU plumbing path)
    ``reduce_scatter`` / ``all_gather`` collectives around the same sharded optimizer.

Both share bucket layout and optimizer state, so checkpoints move freely between them.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Iterable

import torch
import torch.distributed as dist
import torch.nn as nn

from ..ops import _lib, reference
from ..ops.functional import _count
from .mesh import Mesh
from .symm import SymmetricHeap

SHARD_ALIGN = 1024  # elements; also the int8 quantisation block of the outer step


@dataclass
class Bucket:
    name: str
    params: list[tuple[str, nn.Parameter, int]]  # (qualified name, param, offset inside bucket)
    start: int = 0  # offset in the flat buffers
    size: int = 0  # padded element count (multiple of F * SHARD_ALIGN)
    shard_start: int = 0  # offset in this rank's shard buffers
    shard_size: int = 0
    ready: bool = False
    work: object = None  # async collective handle (collective backend)


class _Boundary(torch.autograd.Function):
    """Identity whose backward tells the engine "stage ``bucket_id`` has finished its backward"."""

    @staticmethod
    def forward(ctx, engine, bucket_id, *tensors):
        ctx.engine, ctx.bucket_id = engine, bucket_id
        return tuple(t.view_as(t) for t in tensors)

    @staticmethod
    def backward(ctx, *grads):
        ctx.engine._on_bucket_ready(ctx.bucket_id)
        return (None, None, *grads)


@dataclass
class AdamHyper:
    lr: float = 4e-4
    beta1: float = 0.9
    beta2: float = 0.95
    eps: float = 1e-8
    weight_decay: float = 0.1
    max_norm: float = 1.0


class ShardedEngine:
    def __init__(
        self,
        model: nn.Module,
        mesh: Mesh,
        hyper: AdamHyper,
        *,
        backend: str = "auto",
        heap: SymmetricHeap | None = None,
        overlap: bool = True,
        reduce_ctas: int = 32,
    ):
        self.model, self.mesh, self.hyper = model, mesh, hyper
        self.F = mesh.fsdp_size
        self.device = next(model.parameters()).device
        self.param_dtype = next(model.parameters()).dtype
        if backend == "auto":
            backend = "fused" if (self.device.type == "cuda" and heap is not None) else "collective"
        if backend == "fused" and (heap is None or self.device.type != "cuda"):
            raise ValueError("fused backend needs CUDA and a SymmetricHeap")
        self.backend, self.heap, self.overlap, self.reduce_ctas = backend, heap, overlap, reduce_ctas
        self.step_count = 0
        self.last_grad_norm: torch.Tensor | None = None
        self._last_micro = True
        self._epoch = 0
        self.capture_mode = False  # True while a CUDA graph of the micro-step is captured/replayed: no comm in backward
        self._build_buckets()
        self._allocate()
        self._install_hooks()
        if self.device.type == "cuda":
            self.lib = _lib.load()
            self.comm_stream = torch.cuda.Stream(device=self.device) if overlap else None
        else:
            self.lib, self.comm_stream = None, None

    # ------------------------------------------------------------------ layout
    def _stages(self) -> list[tuple[str, list[tuple[str, nn.Parameter]]]]:
        m = self.model
        named = dict(m.named_parameters())
        used: set[str] = set()
        stages: list[tuple[str, list[tuple[str, nn.Parameter]]]] = []

        def take(prefix_list: Iterable[str], name: str):
            items = [(k, v) for k, v in named.items() if any(k == p or k.startswith(p + ".") for p in prefix_list)]
            used.update(k for k, _ in items)
            stages.append((name, items))

        if hasattr(m, "layers") and hasattr(m, "tok_embeddings"):
            take(["tok_embeddings"], "embed")
            for i in range(len(m.layers)):
                take([f"layers.{i}"], f"layer{i}")
            take(["norm", "output"], "head")
        rest = [(k, v) for k, v in named.items() if k not in used]
        if rest:
            stages.append(("rest", rest))
        return [s for s in stages if s[1]]

    def _build_buckets(self) -> None:
        gran = self.F * SHARD_ALIGN
        self.buckets: list[Bucket] = []
        flat_off = shard_off = 0
        for name, items in self._stages():
            off = 0
            plist = []
            for qn, p in items:
                off = (off + 7) // 8 * 8  # 16-byte alignment of every bf16 parameter (TMA base address)
                plist.append((qn, p, off))
                off += p.numel()
            size = (off + gran - 1) // gran * gran
            b = Bucket(name, plist, flat_off, size, shard_off, size // self.F)
            self.buckets.append(b)
            flat_off += size
            shard_off += b.shard_size
        self.total, self.shard_total = flat_off, shard_off
        self.bucket_index = {b.name: i for i, b in enumerate(self.buckets)}

    def _allocate(self) -> None:
        dev, F, r = self.device, self.F, self.mesh.fsdp_rank
        if self.backend == "fused":
            self.param_flat = self.heap.alloc(self.total, self.param_dtype)
            self.grad_flat = self.heap.alloc(self.total, torch.float32)
            self.slot_grad = self.heap.alloc_flags(len(self.buckets) * F)
            self.slot_norm = self.heap.alloc_flags(F)
            self.slot_bar = self.heap.alloc_flags(F)
            # NVLS reduce-scatter: only when the heap has a multicast mapping AND it spans exactly this FSDP group (the switch
            # sums every device bound to the object, so a heap shared by several DiLoCo workers cannot use it)
            self._nvls = F > 1 and hasattr(self.heap, "mc_ptr") and self.heap.world_size == F
            self.sumsq_partial = torch.zeros(self.lib_grid(), dtype=torch.float32, device=dev)
            self.gnorm_buf = torch.zeros(1, dtype=torch.float32, device=dev)
        else:
            self.param_flat = torch.zeros(self.total, dtype=self.param_dtype, device=dev)
            self.grad_flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.param_flat.zero_()
        self.grad_flat.zero_()
        with torch.no_grad():
            for b in self.buckets:
                for _, p, off in b.params:
                    view = self.param_flat[b.start + off : b.start + off + p.numel()].view(p.shape)
                    view.copy_(p.data)
                    p.data = view
                    p.main_grad = self.grad_flat[b.start + off : b.start + off + p.numel()].view(p.shape)
        # partitioned optimizer state
        self.master = torch.empty(self.shard_total, dtype=torch.float32, device=dev)
        for b in self.buckets:
            s = b.start + r * b.shard_size
            self.master[b.shard_start : b.shard_start + b.shard_size] = self.param_flat[s : s + b.shard_size].float()
        self.exp_avg = torch.zeros_like(self.master)
        self.exp_avg_sq = torch.zeros_like(self.master)
        # reduced gradient shard; with F == 1 the shard *is* the bucket, so alias instead of copying
        self.gshard = self.grad_flat if F == 1 else torch.zeros(self.shard_total, dtype=torch.float32, device=dev)

    def lib_grid(self) -> int:
        return _lib.load().pb_grad_reduce_grid()

    # ------------------------------------------------------------------ hooks
    def _install_hooks(self) -> None:
        m = self.model
        if not (hasattr(m, "layers") and len(m.layers) > 0):
            return
        eng = self

        def pre_hook_for(bucket_id: int):
            def hook(module, args, kwargs):
                if not torch.is_grad_enabled():
                    return None
                idx = [i for i, a in enumerate(args) if isinstance(a, torch.Tensor) and a.requires_grad]
                if not idx:
                    return None
                outs = _Boundary.apply(eng, bucket_id, *[args[i] for i in idx])
                args = list(args)
                for i, o in zip(idx, outs):
                    args[i] = o
                return tuple(args), kwargs

            return hook

        for i, layer in enumerate(m.layers):
            layer.register_forward_pre_hook(pre_hook_for(self.bucket_index[f"layer{i}"]), with_kwargs=True)

        head_id = self.bucket_index.get("head")
        if head_id is not None:

            def post_hook(module, args, output):
                if not torch.is_grad_enabled():
                    return None
                outs = output if isinstance(output, tuple) else (output,)
                idx = [i for i, a in enumerate(outs) if isinstance(a, torch.Tensor) and a.requires_grad]
                if not idx:
                    return None
                wrapped = _Boundary.apply(eng, head_id, *[outs[i] for i in idx])
                outs = list(outs)
                for i, o in zip(idx, wrapped):
                    outs[i] = o
                return tuple(outs) if isinstance(output, tuple) else outs[0]

            m.layers[-1].register_forward_hook(post_hook)

    # ------------------------------------------------------------------ step protocol
    def zero_grad(self) -> None:
        self.grad_flat.zero_()
        for b in self.buckets:
            b.ready, b.work = False, None

    def set_micro_step(self, last: bool) -> None:
        """Tell the engine whether the coming backward is the last of the accumulation window."""
        self._last_micro = last
        if last:
            self._epoch += 1

    def _fold_autograd_grads(self, b: Bucket) -> None:
        """Ops without main_grad fusion (embedding, CPU reference ops) leave ``.grad``: fold into main_grad."""
        for _, p, _ in b.params:
            if p.grad is not None:
                p.main_grad.add_(p.grad.to(torch.float32))
                p.grad = None

    def _on_bucket_ready(self, bucket_id: int) -> None:
        b = self.buckets[bucket_id]
        if self.capture_mode or not self._last_micro or b.ready:
            return
        b.ready = True
        self._fold_autograd_grads(b)
        self._reduce_bucket(bucket_id)

    def _reduce_bucket(self, bucket_id: int) -> None:
        b, F, r = self.buckets[bucket_id], self.F, self.mesh.fsdp_rank
        if self.backend == "fused":
            main = torch.cuda.current_stream()
            cs = self.comm_stream or main
            if cs is not main:
                cs.wait_stream(main)
            with torch.cuda.stream(cs):
                s = cs.cuda_stream
                ranks = self.mesh.fsdp_ranks
                wait_flags = None
                slot_base = self.slot_grad + bucket_id * F
                if F > 1:
                    pp = self.heap.peers(ranks, self.heap.flags)
                    _lib.check(self.lib.pb_signal(ctypes.byref(pp), slot_base + r, self._epoch, s), "pb_signal")
                    _count()
                    wait_flags = self.heap.flags.data_ptr()
                out = self.gshard[b.shard_start :] if F > 1 else self.grad_flat[b.start :]
                if self._nvls:  # the switch sums the F copies: one multimem.ld_reduce per 16 bytes instead of F peer loads
                    _lib.check(
                        self.lib.pb_mc_grad_reduce(
                            self.heap.mc_ptr(self.grad_flat), b.start + r * b.shard_size, b.shard_size, 1.0 / F, out.data_ptr(),
                            self.sumsq_partial.data_ptr(), wait_flags, slot_base, self._epoch, F, self.heap.err.data_ptr(),
                            self.reduce_ctas if self.overlap else 0, s,
                        ),
                        "pb_mc_grad_reduce",
                    )  # fmt: skip
                    _count()
                    return
                gp = self.heap.peers(ranks, self.grad_flat)
                _lib.check(
                    self.lib.pb_grad_reduce(
                        ctypes.byref(gp), b.start + r * b.shard_size, b.shard_size, 1.0 / F, out.data_ptr(),
                        self.sumsq_partial.data_ptr(), wait_flags, slot_base, self._epoch, self.heap.err.data_ptr(),
                        self.reduce_ctas if (F > 1 and self.overlap) else 0, s,
                    ),
                    "pb_grad_reduce",
                )  # fmt: skip
                _count()
            return
        # ---- collective backend
        if F == 1:
            return
        src = self.grad_flat[b.start : b.start + b.size]
        dst = self.gshard[b.shard_start : b.shard_start + b.shard_size]
        g = self.mesh.fsdp_group
        if dist.get_backend(g) == "nccl":
            b.work = dist.reduce_scatter_tensor(dst, src, op=dist.ReduceOp.AVG, group=g, async_op=True)
        else:  # gloo has no reduce_scatter: all-reduce the bucket, keep my slice
            dist.all_reduce(src, group=g)
            dst.copy_(src[r * b.shard_size : (r + 1) * b.shard_size])
            dst.mul_(1.0 / F)

    def finish_backward(self) -> None:
        """Call after ``loss.backward()`` of the last micro-step: reduces whatever has no boundary (embeddings)."""
        for i, b in enumerate(self.buckets):
            if not b.ready:
                b.ready = True
                self._fold_autograd_grads(b)
                self._reduce_bucket(i)

    def fold_micro_grads(self) -> None:
        """After a non-final micro-step backward: fold stray ``.grad`` tensors into main_grad."""
        for b in self.buckets:
            self._fold_autograd_grads(b)

    # ------------------------------------------------------------------ optimizer
    def step(self, lr: float | None = None) -> None:
        h = self.hyper
        lr = h.lr if lr is None else lr
        self.step_count += 1
        t = self.step_count
        bc1, bc2 = 1.0 - h.beta1**t, 1.0 - h.beta2**t
        if self.backend == "fused":
            self._step_fused(lr, bc1, bc2)
        else:
            self._step_collective(lr, bc1, bc2)
        for b in self.buckets:
            b.ready, b.work = False, None

    def _step_fused(self, lr: float, bc1: float, bc2: float) -> None:
        h, F, r = self.hyper, self.F, self.mesh.fsdp_rank
        main = torch.cuda.current_stream()
        cs = self.comm_stream or main
        if cs is not main:
            cs.wait_stream(main)
        ranks = self.mesh.fsdp_ranks
        heap = self.heap
        with torch.cuda.stream(cs):
            s = cs.cuda_stream
            flags_pp = heap.peers(ranks, heap.flags)
            norm_pp = heap.peers(ranks, heap.norm_slots)
            _lib.check(
                self.lib.pb_norm_publish(self.sumsq_partial.data_ptr(), self.sumsq_partial.numel(), ctypes.byref(norm_pp),
                                         ctypes.byref(flags_pp), r, self.slot_norm, self._epoch, s),
                "pb_norm_publish",
            )  # fmt: skip
            _count(1 + len(self.buckets) + (1 if F > 1 else 0))
            args = _lib.AdamArgs(lr, h.beta1, h.beta2, h.eps, h.weight_decay, bc1, bc2, h.max_norm)
            for b in self.buckets:
                dst = heap.peers(ranks, self.param_flat)
                sl = slice(b.shard_start, b.shard_start + b.shard_size)
                g = self.gshard[b.shard_start :] if F > 1 else self.grad_flat[b.start :]
                _lib.check(
                    self.lib.pb_adamw_push(
                        self.master[sl].data_ptr(), g.data_ptr(), self.exp_avg[sl].data_ptr(), self.exp_avg_sq[sl].data_ptr(),
                        b.shard_size, ctypes.byref(args), heap.norm_slots.data_ptr(), F, heap.flags.data_ptr(),
                        self.slot_norm, self._epoch, ctypes.byref(dst), b.start + r * b.shard_size,
                        self.gnorm_buf.data_ptr(), heap.err.data_ptr(), s,
                    ),
                    "pb_adamw_push",
                )  # fmt: skip
            if F > 1:
                heap.barrier(ranks, self.slot_bar, self._epoch, s)
        if cs is not main:
            main.wait_stream(cs)
        self.last_grad_norm = self.gnorm_buf

    def _step_collective(self, lr: float, bc1: float, bc2: float) -> None:
        h, F, r = self.hyper, self.F, self.mesh.fsdp_rank
        g = self.mesh.fsdp_group
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()
        gsh = self.gshard  # aliases grad_flat when F == 1 (identical layout)
        sumsq = gsh.pow(2).sum()
        if F > 1:
            dist.all_reduce(sumsq, group=g)
        gnorm = sumsq.sqrt()
        self.last_grad_norm = gnorm
        clip = torch.clamp(h.max_norm / (gnorm + 1e-6), max=1.0) if h.max_norm > 0 else torch.ones((), device=gnorm.device)
        gsh = gsh * clip
        reference.adamw_step(
            self.master, gsh, self.exp_avg, self.exp_avg_sq, lr=lr, beta1=h.beta1, beta2=h.beta2, eps=h.eps,
            weight_decay=h.weight_decay, step=self.step_count,
        )  # fmt: skip
        self.publish_params()

    @torch.no_grad()
    def publish_params(self) -> None:
        """fp32 master shard → parameter buffers of the whole FSDP group (cast + all-gather)."""
        F, r = self.F, self.mesh.fsdp_rank
        if self.backend == "fused":
            main = torch.cuda.current_stream()
            ranks = self.mesh.fsdp_ranks
            dst = self.heap.peers(ranks, self.param_flat)
            for b in self.buckets:
                sl = slice(b.shard_start, b.shard_start + b.shard_size)
                _lib.check(
                    self.lib.pb_cast_push(self.master[sl].data_ptr(), b.shard_size, ctypes.byref(dst),
                                          b.start + r * b.shard_size, main.cuda_stream),
                    "pb_cast_push",
                )  # fmt: skip
            if F > 1:
                self._epoch += 1
                self.heap.barrier(ranks, self.slot_bar, self._epoch, main.cuda_stream)
            return
        for b in self.buckets:
            mine = self.param_flat[b.start + r * b.shard_size : b.start + (r + 1) * b.shard_size]
            mine.copy_(self.master[b.shard_start : b.shard_start + b.shard_size])
            if F > 1:
                dist.all_gather_into_tensor(self.param_flat[b.start : b.start + b.size], mine.clone(), group=self.mesh.fsdp_group)

    # ------------------------------------------------------------------ state
    def state_dict(self) -> dict:
        return {
            "step": self.step_count,
            "master": self.master,
            "exp_avg": self.exp_avg,
            "exp_avg_sq": self.exp_avg_sq,
            "layout": [(b.name, b.start, b.size, b.shard_start, b.shard_size) for b in self.buckets],
            "fsdp_size": self.F,
            "fsdp_rank": self.mesh.fsdp_rank,
        }

    def load_state_dict(self, sd: dict) -> None:
        if sd["fsdp_size"] != self.F:
            raise ValueError(f"checkpoint was written with fsdp_size={sd['fsdp_size']}, engine has {self.F}")
        self.step_count = int(sd["step"])
        self.master.copy_(sd["master"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.publish_params()

    def param_hash(self) -> float:
        """Cheap replica-consistency probe (sum of the bf16 parameter buffer)."""
        return float(self.param_flat.float().sum().item())
