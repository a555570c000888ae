"""DiLoCo outer optimizer: every H inner steps, all-reduce the (int8-quantised) pseudo-gradient
across workers and take a Nesterov-SGD step on the replicated outer parameters θ₀.

State lives on the GPU (θ₀ and the momentum buffer are sharded exactly like the inner optimizer:
1/F per rank, so 180 GB parts never need the host offload the H100-era engine used).

Exchange design: int8 payloads cannot be summed in the switch, so the all-reduce is
an *all-gather of (int8, per-1024 scales)* followed by a local dequantise-and-sum in fixed worker
order — every worker ends up with bitwise-identical θ₀ with one quantisation per contribution.

``backend="fused"``  : kernel 1 = pseudo-gradient ⊕ quantise (into the symmetric heap);
                       flag barrier; kernel 2 = peer int8 loads ⊕ dequant-sum ⊕ Nesterov ⊕ inner
                       master reset ⊕ bf16 parameter stores to the FSDP group.
``backend="collective"``: same math with ``all_gather`` / ``all_reduce`` (NCCL or gloo) — baseline + CPU path.
"""

from __future__ import annotations

import ctypes
import time
from dataclasses import dataclass

import torch
import torch.distributed as dist

from ..ops import _lib, reference
from ..ops.functional import _count
from .fsdp import SHARD_ALIGN, ShardedEngine


@dataclass
class OuterHyper:
    lr: float = 0.7
    momentum: float = 0.9
    nesterov: bool = True
    compression: str = "int8"  # "int8" | "no"


def _all_gather(group, outs: list[torch.Tensor], t: torch.Tensor) -> None:
    """Works for registered groups and for the standalone per-epoch groups the elastic coordinator builds."""
    if group is None:
        dist.all_gather(outs, t)
    else:
        group.allgather([outs], [t]).wait()


def _all_reduce(group, t: torch.Tensor) -> None:
    if group is None:
        dist.all_reduce(t)
    else:
        group.allreduce([t]).wait()


class DilocoOuter:
    def __init__(self, engine: ShardedEngine, hyper: OuterHyper, *, diloco_group=None, diloco_ranks=None, collective: bool = False):
        self.engine, self.hyper = engine, hyper
        self.mesh = engine.mesh
        self.group = diloco_group if diloco_group is not None else self.mesh.diloco_group
        self.ranks = list(diloco_ranks) if diloco_ranks is not None else list(self.mesh.diloco_ranks)
        n = engine.shard_total
        dev = engine.device
        self.theta0 = engine.master.clone()
        self.momentum = torch.zeros(n, dtype=torch.float32, device=dev)
        self.outer_step_count = 0
        self.last_bytes_on_wire = 0
        self.last_seconds = 0.0
        # ``collective=True``: elastic jobs — workers are separate process worlds, so the exchange goes through a
        # (re-creatable) process group instead of the symmetric heap, which cannot span independently launched workers
        self.fused = engine.backend == "fused" and not collective
        if self.fused:
            heap = engine.heap
            self.q = heap.alloc(n, torch.int8)
            self.scales = heap.alloc(n // SHARD_ALIGN, torch.float32)
            world = heap.world_size
            self.slot_bar = heap.alloc_flags(world)
            self.slot_bar2 = heap.alloc_flags(world)
            self._epoch = 0

    @property
    def num_workers(self) -> int:
        return len(self.ranks)

    def set_membership(self, ranks, group) -> None:
        """Elastic join/drop: swap the set of peers contributing to the next outer step."""
        self.ranks, self.group = list(ranks), group

    # ------------------------------------------------------------------ the outer step
    @torch.no_grad()
    def step(self) -> None:
        t0 = time.perf_counter()
        if self.fused:
            self._step_fused()
        else:
            self._step_collective()
        self.outer_step_count += 1
        self.last_seconds = time.perf_counter() - t0

    def _step_fused(self) -> None:
        eng, heap, lib = self.engine, self.engine.heap, self.engine.lib
        s = torch.cuda.current_stream().cuda_stream
        n = eng.shard_total
        W = self.num_workers
        all_ranks = list(range(heap.world_size))
        if self.hyper.compression == "int8":
            _lib.check(lib.pb_pseudograd_quant(self.theta0.data_ptr(), eng.master.data_ptr(), self.q.data_ptr(),
                                               self.scales.data_ptr(), n, s), "pb_pseudograd_quant")  # fmt: skip
            self._epoch += 1
            heap.barrier(all_ranks, self.slot_bar, self._epoch, s)  # everyone's payload is published
            args = _lib.OuterArgs(self.hyper.lr, self.hyper.momentum, 1.0 / W, int(self.hyper.nesterov))
            dst = heap.peers(self.mesh.fsdp_ranks, eng.param_flat)
            r = self.mesh.fsdp_rank
            for b in eng.buckets:
                lo = b.shard_start
                qs = _lib.PeerPtrs.of(heap.peer_ptr(w, self.q) + lo for w in self.ranks)
                ss = _lib.PeerPtrs.of(heap.peer_ptr(w, self.scales) + (lo // SHARD_ALIGN) * 4 for w in self.ranks)
                _lib.check(
                    lib.pb_outer_nesterov(ctypes.byref(qs), ctypes.byref(ss), self.theta0[lo:].data_ptr(),
                                          self.momentum[lo:].data_ptr(), eng.master[lo:].data_ptr(), b.shard_size,
                                          ctypes.byref(args), ctypes.byref(dst), b.start + r * b.shard_size, s),
                    "pb_outer_nesterov",
                )  # fmt: skip
            self._epoch += 1
            heap.barrier(all_ranks, self.slot_bar2, self._epoch, s)  # payloads consumed, params landed
            self.last_bytes_on_wire = (W - 1) * (n + 4 * (n // SHARD_ALIGN))
            _count(3 + len(eng.buckets))
        else:
            self._epoch += 1
            heap.barrier(all_ranks, self.slot_bar, self._epoch, s)
            # uncompressed: peers read each other's fp32 inner masters directly (they must live in the heap)
            raise NotImplementedError("fused fp32 outer path requires heap-resident masters; use compression='int8'")

    def _step_collective(self) -> None:
        eng, h = self.engine, self.hyper
        W = self.num_workers
        pseudo = self.theta0 - eng.master
        if W > 1:
            if h.compression == "int8":
                q, sc = reference.quantize_int8_blockwise(pseudo, SHARD_ALIGN)
                qs = [torch.empty_like(q) for _ in range(W)]
                scs = [torch.empty_like(sc) for _ in range(W)]
                _all_gather(self.group, qs, q)
                _all_gather(self.group, scs, sc)
                avg = torch.zeros_like(pseudo)
                for qw, sw in zip(qs, scs):  # fixed worker order → identical result on every worker
                    avg += reference.dequantize_int8_blockwise(qw, sw, SHARD_ALIGN)
                avg /= W
                self.last_bytes_on_wire = (W - 1) * (q.numel() + 4 * sc.numel())
            else:
                avg = pseudo
                _all_reduce(self.group, avg)
                avg /= W
                self.last_bytes_on_wire = 2 * (W - 1) * 4 * avg.numel() // W
        else:
            if h.compression == "int8":
                q, sc = reference.quantize_int8_blockwise(pseudo, SHARD_ALIGN)
                avg = reference.dequantize_int8_blockwise(q, sc, SHARD_ALIGN)
            else:
                avg = pseudo
            self.last_bytes_on_wire = 0
        reference.nesterov_outer_step(self.theta0, avg, self.momentum, lr=h.lr, momentum=h.momentum, nesterov=h.nesterov)
        eng.master.copy_(self.theta0)
        eng.publish_params()

    # ------------------------------------------------------------------ state
    def state_dict(self) -> dict:
        return {"theta0": self.theta0, "momentum": self.momentum, "outer_step": self.outer_step_count}

    def load_state_dict(self, sd: dict) -> None:
        self.theta0.copy_(sd["theta0"])
        self.momentum.copy_(sd["momentum"])
        self.outer_step_count = int(sd["outer_step"])
