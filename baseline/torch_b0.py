"""B0 — the comparison baseline: the SAME job (Llama, DiLoCo H, two-level mesh, batch plan, int8 outer exchange) written with
stock PyTorch only.  Nothing from ``prime_b200`` is imported here: no native kernel, no engine, no symmetric heap.

    model        nn.Embedding / nn.RMSNorm / nn.Linear (cuBLAS) / F.scaled_dot_product_attention (cuDNN / flash) / F.silu,
                 fused QKV and gate-up projections (the generous choice: fewer, larger cuBLAS GEMMs than torchtitan's split ones)
    inner loop   FSDP2 ``fully_shard`` per block over the worker's FSDP sub-mesh (NCCL all-gather / reduce-scatter, bf16 compute,
                 fp32 reduce, fp32 sharded master weights), gradient accumulation with ``set_requires_gradient_sync``,
                 ``clip_grad_norm_`` and fused ``torch.optim.AdamW``
    outer step   θ₀ − θ on the local shards → int8 block quantisation (1024) → ``dist.all_gather`` over the DiLoCo group →
                 dequantise-sum → Nesterov SGD on θ₀ → copy back into the sharded parameters
    data         the same synthetic token stream shape; every micro-batch is copied from pinned host memory

``bench.py`` runs this arm back to back with the engine in the same invocation and prints the ratio (``vs_b0``); it is also
the numerical oracle of the loss-curve comparison (``tools/loss_curve.py``).  SURVEY.md §7.1 step 2 / BASELINE.md "What the
build must do instead" define it; the mounted reference itself contains no trainer.
"""

from __future__ import annotations

import math
import os
import time
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

SIZES = {
    "debugmodel": dict(dim=256, n_layers=2, n_heads=8, vocab=2048),
    "150M": dict(dim=1024, n_layers=12, n_heads=16, vocab=32000),
    "1B": dict(dim=2048, n_layers=18, n_heads=16, vocab=32000),
    "7B": dict(dim=4096, n_layers=32, n_heads=32, vocab=32000),
}


def ffn_hidden(dim: int, multiple_of: int = 256) -> int:
    h = int(2 * (4 * dim) / 3)
    return multiple_of * ((h + multiple_of - 1) // multiple_of)


def rope_freqs(seq: int, head_dim: int, theta: float, device) -> torch.Tensor:
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, device=device, dtype=torch.float32) / head_dim))
    ang = torch.outer(torch.arange(seq, device=device, dtype=torch.float32), inv)
    return torch.polar(torch.ones_like(ang), ang)  # complex64 [S, D/2]


def apply_rope(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    # x [B, S, H, D] with interleaved pairs (the Llama reference formulation)
    xc = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    out = torch.view_as_real(xc * freqs[None, : x.shape[1], None, :]).flatten(3)
    return out.type_as(x)


class Block(nn.Module):
    def __init__(self, dim: int, n_heads: int, layer_id: int, n_layers: int):
        super().__init__()
        self.n_heads, self.head_dim = n_heads, dim // n_heads
        hid = ffn_hidden(dim)
        self.attention_norm = nn.RMSNorm(dim, eps=1e-5)
        self.ffn_norm = nn.RMSNorm(dim, eps=1e-5)
        self.wqkv = nn.Linear(dim, 3 * dim, bias=False)
        self.wo = nn.Linear(dim, dim, bias=False)
        self.w13 = nn.Linear(dim, 2 * hid, bias=False)
        self.w2 = nn.Linear(hid, dim, bias=False)
        std = 0.02 / math.sqrt(2 * (layer_id + 1))
        nn.init.normal_(self.wqkv.weight, 0.0, 0.02)
        nn.init.normal_(self.w13.weight, 0.0, 0.02)
        nn.init.normal_(self.wo.weight, 0.0, std)
        nn.init.normal_(self.w2.weight, 0.0, std)

    def forward(self, x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
        B, S, _ = x.shape
        q, k, v = self.wqkv(self.attention_norm(x)).view(B, S, 3, self.n_heads, self.head_dim).unbind(2)
        q, k = apply_rope(q, freqs), apply_rope(k, freqs)
        a = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True)
        x = x + self.wo(a.transpose(1, 2).reshape(B, S, -1))
        g, u = self.w13(self.ffn_norm(x)).chunk(2, dim=-1)
        return x + self.w2(F.silu(g) * u)


class Llama(nn.Module):
    def __init__(self, dim: int, n_layers: int, n_heads: int, vocab: int, max_seq: int, rope_theta: float = 10000.0):
        super().__init__()
        self.tok_embeddings = nn.Embedding(vocab, dim)
        self.layers = nn.ModuleList(Block(dim, n_heads, i, n_layers) for i in range(n_layers))
        self.norm = nn.RMSNorm(dim, eps=1e-5)
        self.output = nn.Linear(dim, vocab, bias=False)
        nn.init.normal_(self.tok_embeddings.weight, 0.0, 1.0)
        nn.init.trunc_normal_(self.output.weight, 0.0, dim**-0.5, -3 * dim**-0.5, 3 * dim**-0.5)
        self.head_dim, self.rope_theta, self.max_seq = dim // n_heads, rope_theta, max_seq
        self._freqs: torch.Tensor | None = None

    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        if self._freqs is None or self._freqs.device != tokens.device:
            self._freqs = rope_freqs(self.max_seq, self.head_dim, self.rope_theta, tokens.device)
        h = self.tok_embeddings(tokens)
        for layer in self.layers:
            h = layer(h, self._freqs)
        return self.output(self.norm(h))

    def flops_per_token(self, seq: int) -> float:
        n_mm = sum(p.numel() for p in self.parameters()) - self.tok_embeddings.weight.numel()
        dim, L = self.tok_embeddings.weight.shape[1], len(self.layers)
        return 6.0 * n_mm + 6.0 * L * seq * dim


# ------------------------------------------------------------------------------------------------- int8 block quantisation
QBLOCK = 1024


def quantize_int8(x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    n = x.numel()
    pad = (-n) % QBLOCK
    xp = F.pad(x.reshape(-1), (0, pad)).view(-1, QBLOCK)
    scale = xp.abs().amax(dim=1) / 127.0
    inv = torch.where(scale > 0, 1.0 / scale, torch.zeros_like(scale))
    q = torch.clamp(torch.round(xp * inv[:, None]), -127, 127).to(torch.int8)
    return q, scale


def dequantize_int8(q: torch.Tensor, scale: torch.Tensor, n: int) -> torch.Tensor:
    return (q.float() * scale[:, None]).reshape(-1)[:n]


# ------------------------------------------------------------------------------------------------- data
class PinnedFakeTokens:
    """Synthetic next-token stream; every batch is produced on the host in pinned memory and copied to the device."""

    def __init__(self, vocab: int, seq: int, batch: int, device: torch.device, seed: int):
        self.rng = np.random.default_rng(seed)
        self.vocab, self.seq, self.batch, self.device = vocab, seq, batch, device
        self.host = [torch.empty((2, batch, seq), dtype=torch.int64, pin_memory=device.type == "cuda") for _ in range(4)]
        self.done = [None] * 4
        self.i = 0
        self.h2d_bytes_per_batch = 2 * batch * seq * 8

    def next(self) -> tuple[torch.Tensor, torch.Tensor]:
        k = self.i % 4
        self.i += 1
        if self.done[k] is not None:
            self.done[k].synchronize()
        start = self.rng.integers(0, self.vocab, size=(self.batch, 1), dtype=np.int64)
        steps = self.rng.integers(1, 4, size=(self.batch, self.seq), dtype=np.int64)
        toks = (start + np.cumsum(steps, axis=1)) % self.vocab
        full = np.concatenate([start % self.vocab, toks], axis=1)
        self.host[k][0].copy_(torch.from_numpy(np.ascontiguousarray(full[:, :-1])))
        self.host[k][1].copy_(torch.from_numpy(np.ascontiguousarray(full[:, 1:])))
        dev = self.host[k].to(self.device, non_blocking=True)
        if self.device.type == "cuda":
            self.done[k] = torch.cuda.Event()
            self.done[k].record()
        return dev[0], dev[1]


# ------------------------------------------------------------------------------------------------- trainer
@dataclass
class B0Config:
    model: str = "1B"
    seq: int = 1024
    micro_bs: int = 16
    accum: int = 4
    workers: int = 1
    fsdp: int = 1
    inner_steps: int = 100
    lr: float = 4e-4
    outer_lr: float = 0.7
    outer_momentum: float = 0.9
    compression: str = "int8"
    max_norm: float = 1.0
    seed: int = 42
    compile: bool = False  # torch.compile of each block (needs a working inductor/triton toolchain on the box)
    diloco: bool = True


class B0Trainer:
    def __init__(self, cfg: B0Config):
        from torch.distributed.device_mesh import init_device_mesh
        from torch.distributed.fsdp import MixedPrecisionPolicy, fully_shard

        self.cfg = cfg
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        local = int(os.environ.get("LOCAL_RANK", 0))
        cuda = torch.cuda.is_available()
        self.device = torch.device("cuda", local) if cuda else torch.device("cpu")
        if cuda:
            torch.cuda.set_device(local)
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl" if cuda else "gloo", rank=self.rank, world_size=self.world,
                                    **({"device_id": self.device} if cuda else {}))  # fmt: skip
        assert cfg.workers * cfg.fsdp == self.world
        self.mesh = init_device_mesh(self.device.type, (cfg.workers, cfg.fsdp), mesh_dim_names=("diloco", "fsdp"))
        self.fsdp_mesh = self.mesh["fsdp"]
        self.diloco_group = self.mesh["diloco"].get_group() if cfg.workers > 1 else None
        torch.manual_seed(cfg.seed)
        sz = SIZES[cfg.model]
        with torch.device(self.device):
            self.model = Llama(sz["dim"], sz["n_layers"], sz["n_heads"], sz["vocab"], max_seq=max(cfg.seq, 128))
        self.vocab = sz["vocab"]
        mp = MixedPrecisionPolicy(param_dtype=torch.bfloat16, reduce_dtype=torch.float32) if cuda else MixedPrecisionPolicy()
        for blk in self.model.layers:
            fully_shard(blk, mesh=self.fsdp_mesh, mp_policy=mp)
        fully_shard(self.model, mesh=self.fsdp_mesh, mp_policy=mp)
        if cfg.compile:
            for blk in self.model.layers:
                blk.compile()
        self.opt = torch.optim.AdamW(self.model.parameters(), lr=cfg.lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1, fused=cuda)
        self.loader = PinnedFakeTokens(self.vocab, cfg.seq, cfg.micro_bs, self.device, seed=cfg.seed * 10007 + self.rank)
        self.step_count = 0
        self.tokens_per_step = cfg.micro_bs * cfg.accum * cfg.seq * self.world
        self.outer_seconds: list[float] = []
        self.outer_bytes = 0
        if cfg.diloco:
            locals_ = [self._local(p) for p in self.model.parameters()]
            self.theta0 = [t.detach().clone().float() for t in locals_]
            self.outer_opt = torch.optim.SGD(self.theta0, lr=cfg.outer_lr, momentum=cfg.outer_momentum, nesterov=True)

    @staticmethod
    def _local(p: torch.Tensor) -> torch.Tensor:
        return p.to_local() if hasattr(p, "to_local") else p

    def flops_per_step(self) -> float:
        return self.model.flops_per_token(self.cfg.seq) * self.tokens_per_step

    def inner_step(self) -> torch.Tensor:
        cfg = self.cfg
        self.opt.zero_grad(set_to_none=True)
        loss_acc = torch.zeros((), device=self.device)
        for i in range(cfg.accum):
            x, y = self.loader.next()
            last = i == cfg.accum - 1
            self.model.set_requires_gradient_sync(last)
            logits = self.model(x)
            loss = F.cross_entropy(logits.float().view(-1, logits.shape[-1]), y.reshape(-1)) / cfg.accum
            loss.backward()
            loss_acc += loss.detach()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), cfg.max_norm)
        warm = min(1.0, (self.step_count + 1) / 10)
        for g in self.opt.param_groups:
            g["lr"] = cfg.lr * warm
        self.opt.step()
        self.step_count += 1
        if cfg.diloco and self.step_count % cfg.inner_steps == 0:
            self.outer_step()
        return loss_acc

    @torch.no_grad()
    def outer_step(self) -> None:
        cfg = self.cfg
        W = cfg.workers
        cuda = self.device.type == "cuda"
        if cuda:
            if self.world > 1:
                dist.barrier()  # line the ranks up: arrival skew between workers is not part of the exchange time (same as the engine arm)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        params = [self._local(p) for p in self.model.parameters()]
        sizes = [t.numel() for t in params]
        pseudo = torch.cat([(t0 - p.float()).reshape(-1) for t0, p in zip(self.theta0, params)])
        n = pseudo.numel()
        if cfg.compression == "int8":
            q, sc = quantize_int8(pseudo)
            if W > 1:
                qs = [torch.empty_like(q) for _ in range(W)]
                scs = [torch.empty_like(sc) for _ in range(W)]
                dist.all_gather(qs, q, group=self.diloco_group)
                dist.all_gather(scs, sc, group=self.diloco_group)
                self.outer_bytes = (W - 1) * (q.numel() + 4 * sc.numel())
            else:
                qs, scs = [q], [sc]
                self.outer_bytes = 0
            avg = torch.zeros(n, device=self.device)
            for qw, sw in zip(qs, scs):
                avg += dequantize_int8(qw, sw, n)
            avg /= W
        else:
            avg = pseudo
            if W > 1:
                dist.all_reduce(avg, group=self.diloco_group)
                avg /= W
            self.outer_bytes = 2 * (W - 1) * 4 * n // max(W, 1)
        off = 0
        for t0, k in zip(self.theta0, sizes):
            t0.grad = avg[off : off + k].view_as(t0)
            off += k
        self.outer_opt.step()
        for t0, p in zip(self.theta0, params):
            p.copy_(t0)
            t0.grad = None
        if cuda:
            e1.record()
            e1.synchronize()
            self.outer_seconds.append(e0.elapsed_time(e1) / 1e3)


def run_bench(cfg: B0Config, steps: int, warmup: int) -> dict:
    """Same measurement protocol as bench.py's own arm: W warm-up steps, K device-timed steps (CUDA events, max over ranks) with
    >= 1 outer step inside, then K end-to-end steps with a device→host read of the loss every step."""
    t = B0Trainer(cfg)
    dev, world = t.device, t.world

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def region(n: int, read_loss: bool) -> tuple[float, float]:
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        outer0 = len(t.outer_seconds)
        for i in range(n):
            loss = t.inner_step()
            if i == n - 1 and cfg.diloco and len(t.outer_seconds) == outer0:
                t.outer_step()
            if read_loss:
                v = float(loss.item())
                if v != v:
                    raise RuntimeError("B0 loss is NaN")
        e1.record()
        sync_all()
        host = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), host * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1]) / 1e3

    for _ in range(warmup):
        t.inner_step()
    if cfg.diloco:
        t.outer_step()
    sync_all()
    t.outer_seconds.clear()
    dev_ms, _ = region(steps, False)
    e2e_ms, e2e_host = region(steps, True)
    tokens = t.tokens_per_step * steps
    outer_s = sum(t.outer_seconds) / max(1, len(t.outer_seconds))
    out = {
        "impl": "torch_b0",
        "stack": f"torch {torch.__version__}: FSDP2 fully_shard + NCCL + cuBLAS nn.Linear + SDPA + fused AdamW" + (" + torch.compile" if cfg.compile else ", eager"),
        "value": round(tokens / (dev_ms / 1e3), 1),
        "unit": "tokens/s",
        "ms_per_step": round(dev_ms / steps, 3),
        "e2e": {"value": round(tokens / max(e2e_host, e2e_ms / 1e3), 1), "unit": "tokens/s",
                "h2d_bytes_per_step": t.loader.h2d_bytes_per_batch * cfg.accum, "d2h_bytes_per_step": 4},
        "outer_ms": round(outer_s * 1e3, 3),
        "outer_allreduce_GBps": round(t.outer_bytes / outer_s / 1e9, 2) if outer_s > 0 and t.outer_bytes else None,
        "mfu_flops_per_step": t.flops_per_step(),
    }  # fmt: skip
    del t
    torch.cuda.empty_cache()
    return out
