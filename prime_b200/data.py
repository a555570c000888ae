"""Data pipeline: synthetic token streams (the sandbox has no network) and memory-mapped token files.

Batches are produced on the host in *pinned* memory and copied host→device asynchronously on a side
stream, double-buffered, so the H2D copy of step i+1 overlaps the compute of step i.  The loader is
checkpointable (``state_dict``: RNG state / file cursor).
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np
import torch


@dataclass
class Batch:
    input_ids: torch.Tensor  # [B, S] int64
    labels: torch.Tensor  # [B, S] int64


class FakeTokenDataset:
    """Deterministic pseudo-random tokens with a learnable structure (next token depends on the previous
    one), so the loss measurably decreases in smoke tests."""

    def __init__(self, vocab_size: int, seq_len: int, seed: int = 0, rank: int = 0, world: int = 1):
        self.vocab_size, self.seq_len = vocab_size, seq_len
        self.rng = np.random.default_rng(seed * 10007 + rank)
        self.rank, self.world = rank, world
        self.n_served = 0

    def next_batch(self, batch_size: int) -> tuple[np.ndarray, np.ndarray]:
        V, S = self.vocab_size, self.seq_len
        start = self.rng.integers(0, V, size=(batch_size, 1), dtype=np.int64)
        steps = self.rng.integers(1, 4, size=(batch_size, S), dtype=np.int64)
        toks = (start + np.cumsum(steps, axis=1)) % V
        full = np.concatenate([start % V, toks], axis=1)
        self.n_served += batch_size
        return full[:, :-1], full[:, 1:]

    def state_dict(self) -> dict:
        return {"rng": self.rng.bit_generator.state, "n_served": self.n_served}

    def load_state_dict(self, sd: dict) -> None:
        self.rng.bit_generator.state = sd["rng"]
        self.n_served = sd["n_served"]


class MemmapTokenDataset:
    """Flat uint16/uint32 token file(s); each rank reads a strided set of windows."""

    def __init__(self, paths: str, seq_len: int, rank: int = 0, world: int = 1, dtype=np.uint16):
        files = [Path(p) for p in paths.split(",") if p]
        if not files:
            raise ValueError("no dataset paths given")
        self.arrays = [np.memmap(f, dtype=dtype, mode="r") for f in files]
        self.seq_len, self.rank, self.world = seq_len, rank, world
        self.cursor = 0
        self.windows = sum((len(a) - 1) // seq_len for a in self.arrays)

    def _window(self, idx: int) -> np.ndarray:
        for a in self.arrays:
            n = (len(a) - 1) // self.seq_len
            if idx < n:
                s = idx * self.seq_len
                return np.asarray(a[s : s + self.seq_len + 1], dtype=np.int64)
            idx -= n
        raise IndexError

    def next_batch(self, batch_size: int) -> tuple[np.ndarray, np.ndarray]:
        rows = []
        for _ in range(batch_size):
            idx = (self.cursor * self.world + self.rank) % self.windows
            rows.append(self._window(idx))
            self.cursor += 1
        full = np.stack(rows)
        return full[:, :-1], full[:, 1:]

    def state_dict(self) -> dict:
        return {"cursor": self.cursor}

    def load_state_dict(self, sd: dict) -> None:
        self.cursor = sd["cursor"]


class PinnedPrefetcher:
    """Ring of ``depth`` pinned-host → device slots; the H2D copy of batch i+k overlaps the compute of batch i.

    Contract: a batch returned by :meth:`next` stays valid until the following ``next()`` call *returns* (the trainer
    enqueues every kernel that reads batch i before it asks for batch i+1).  Two hazards are fenced per slot:

    * host side — the pinned buffer of a slot is rewritten on the CPU only after the previous H2D copy out of it has
      finished (``copied[i].synchronize()``); otherwise a host that runs ahead of the GPU (CUDA graphs, log_interval > 1,
      ``read_loss=False`` benches) would let the pending DMA read a newer or torn batch;
    * device side — the copy into a device slot waits for the event recorded after the last consumer of that slot
      (``consumed[i]``), not for everything enqueued on the compute stream, so prefetch really runs ahead.
    """

    def __init__(self, dataset, batch_size: int, device: torch.device, depth: int = 4):
        self.ds, self.bs, self.device = dataset, batch_size, device
        self.cuda = device.type == "cuda"
        S = dataset.seq_len
        self.depth = max(2, depth)
        self.host = [torch.empty((2, batch_size, S), dtype=torch.int64, pin_memory=self.cuda) for _ in range(self.depth)]
        if self.cuda:
            self.dev = [torch.empty((2, batch_size, S), dtype=torch.int64, device=device) for _ in range(self.depth)]
            self.stream = torch.cuda.Stream(device=device)
            self.copied = [torch.cuda.Event() for _ in range(self.depth)]  # H2D out of host[i] / into dev[i] finished
            self.consumed = [torch.cuda.Event() for _ in range(self.depth)]  # every kernel reading dev[i] was enqueued before this
            self._copied_armed = [False] * self.depth
            self._consumed_armed = [False] * self.depth
        self.slot = 0
        self._handed: int | None = None  # slot currently owned by the consumer
        self.h2d_bytes_per_batch = 2 * batch_size * S * 8
        self.host_waits = 0  # times the CPU had to wait for an in-flight DMA before reusing a pinned slot (diagnostic)
        self._inflight: list[int] = []
        self._states: list[dict] = []  # dataset state BEFORE each in-flight batch was drawn (checkpoint = oldest one)
        for _ in range(self.depth - 1):
            self._issue()

    def state_dict(self) -> dict:
        """Position of the next batch the *consumer* will see (prefetched-but-unconsumed batches are replayed on resume)."""
        import copy

        return copy.deepcopy(self._states[0]) if self._states else self.ds.state_dict()

    def load_state_dict(self, sd: dict) -> None:
        self.ds.load_state_dict(sd)
        self.reset()

    def _issue(self) -> None:
        i = self.slot
        import copy

        self._states.append(copy.deepcopy(self.ds.state_dict()))
        x, y = self.ds.next_batch(self.bs)
        if self.cuda and self._copied_armed[i]:
            if not self.copied[i].query():
                self.host_waits += 1
            self.copied[i].synchronize()  # the DMA that last read host[i] is done: safe to rewrite it on the CPU
        self.host[i][0].copy_(torch.from_numpy(np.ascontiguousarray(x)))
        self.host[i][1].copy_(torch.from_numpy(np.ascontiguousarray(y)))
        if self.cuda:
            if self._consumed_armed[i]:
                self.stream.wait_event(self.consumed[i])  # the last consumer of dev[i] is done before we overwrite it
            with torch.cuda.stream(self.stream):
                self.dev[i].copy_(self.host[i], non_blocking=True)
                self.copied[i].record(self.stream)
            self._copied_armed[i] = True
        self._inflight.append(i)
        self.slot = (self.slot + 1) % self.depth

    def _release_handed(self) -> None:
        if self.cuda and self._handed is not None:
            self.consumed[self._handed].record(torch.cuda.current_stream())
            self._consumed_armed[self._handed] = True
        self._handed = None

    def reset(self) -> None:
        """Drop prefetched batches (they were drawn before a dataset ``load_state_dict``) and refill the pipeline."""
        if self.cuda:
            torch.cuda.current_stream().synchronize()
            self.stream.synchronize()
            self._copied_armed = [False] * self.depth
            self._consumed_armed = [False] * self.depth
        self._handed = None
        self._inflight.clear()
        self._states.clear()
        self.slot = 0
        for _ in range(self.depth - 1):
            self._issue()

    def next(self) -> Batch:
        self._release_handed()  # the batch handed out last time has been fully enqueued by now
        self._issue()
        i = self._inflight.pop(0)
        self._states.pop(0)
        if self.cuda:
            torch.cuda.current_stream().wait_event(self.copied[i])
            t = self.dev[i]
            self._handed = i
        else:
            t = self.host[i].clone()
        return Batch(t[0], t[1])


def build_dataset(cfg_data, vocab_size: int, rank: int, world: int):
    if cfg_data.fake or not cfg_data.dataset_name_or_paths:
        return FakeTokenDataset(vocab_size, cfg_data.seq_length, cfg_data.seed, rank, world)
    return MemmapTokenDataset(cfg_data.dataset_name_or_paths, cfg_data.seq_length, rank, world)
