"""NVLink exchange region for the fused outer step of ELASTIC workers.

Elastic workers are separately launched process worlds, so the symmetric heap of a worker (built over its own
``torch.distributed`` world) cannot span them, and round 1 fell back to a per-epoch NCCL group for the int8 pseudo-gradient
exchange — exactly in the configuration (BASELINE config 5, 4→3→4 workers) where the fused kernel matters.  This module gives
every rank a small device region

    [ flags: 64 × uint32 | error word: 16 × uint32 | int8 payload: n | fp32 scales: n / 1024 ]

exported with ``cudaIpcGetMemHandle`` and published through the job's global ``TCPStore`` under ``xchg/<worker>/<fsdp_rank>``.
At every membership change a rank opens the regions of the SAME fsdp_rank in the other member workers (handles are cached per
worker incarnation and closed when a worker is dropped).  The outer step then runs the same kernel pair as the static mesh —
``pseudograd_quant`` → flag barrier → ``outer_nesterov`` (peer int8 loads ⊕ dequant-sum ⊕ Nesterov ⊕ bf16 write-back) → flag
barrier — with the peer pointers of the current members.

Failure model: the device-side waits are bounded (``pb_set_spin_timeout_ms``, a fraction of the heartbeat timeout); a barrier
that times out sets the region's error word, ``outer_nesterov`` sees it and leaves θ₀ / momentum / master untouched, and
``finish()`` raises so that ``train._elastic_setup.guarded_step`` re-runs the rendezvous and retries on the re-formed group.
All GPUs of the box must be visible to every worker (``launch.py`` passes ``PRIME_B200_DEVICES`` instead of slicing
``CUDA_VISIBLE_DEVICES``).

Barrier epochs are ``(membership epoch << 20) + sequence`` (compared wrap-safe on the device): the membership epoch grows at
every rendezvous for every member, so a value left in a flag by an earlier membership (or by a dead worker's previous
incarnation) can never satisfy a later wait — as long as one membership lasts fewer than 2^19 outer steps (two barriers each).
"""

from __future__ import annotations

import ctypes
import json

import torch

from ..ops import _lib
from .symm import _DevBuffer

FLAG_WORDS = 64
ERR_WORDS = 16
SLOT_A, SLOT_B = 0, 16  # first flag slot of the two barriers (≤ 16 workers per box)
QBLOCK = 1024
SEQ_BITS = 20  # barrier sequence numbers per membership epoch (see the module docstring)


class ElasticExchange:
    def __init__(self, store, wid: str, fsdp_rank: int, n_elems: int, device: torch.device):
        assert n_elems % QBLOCK == 0
        self.lib = _lib.load()
        self.store, self.wid, self.r, self.n, self.device = store, wid, fsdp_rank, n_elems, device
        head = (FLAG_WORDS + ERR_WORDS) * 4
        self.q_off = (head + 1023) // 1024 * 1024
        self.s_off = (self.q_off + n_elems + 1023) // 1024 * 1024
        self.nbytes = ((self.s_off + (n_elems // QBLOCK) * 4 + (1 << 21) - 1) >> 21) << 21
        base = ctypes.c_void_p()
        _lib.check(self.lib.pb_ipc_alloc(ctypes.byref(base), self.nbytes), "pb_ipc_alloc")
        self.base = int(base.value)
        whole = torch.as_tensor(_DevBuffer(self.base, self.nbytes), device=device)
        self.flags = whole[: FLAG_WORDS * 4].view(torch.int32)
        self.err = whole[FLAG_WORDS * 4 : head].view(torch.int32)
        self.q = whole[self.q_off : self.q_off + n_elems].view(torch.int8)
        self.scales = whole[self.s_off : self.s_off + (n_elems // QBLOCK) * 4].view(torch.float32)
        self._whole = whole
        hbuf = ctypes.create_string_buffer(self.lib.pb_ipc_handle_size())
        _lib.check(self.lib.pb_ipc_get_handle(self.base, hbuf), "pb_ipc_get_handle")
        torch.cuda.synchronize(device)  # pb_ipc_alloc's clearing memset is asynchronous: finish it before any peer can map and write flags
        self.store.set(self._key(wid), json.dumps({"handle": bytes(hbuf.raw).hex(), "nbytes": self.nbytes, "n": n_elems,
                                                   "device": torch.cuda.current_device()}))  # fmt: skip
        self._open: dict[str, int] = {}  # worker id (with incarnation) → mapped base of its region for my fsdp_rank
        self.members: list[str] = [wid]
        self.index = 0
        self._mem_epoch = 0
        self._seq = 0

    def _key(self, wid: str) -> str:
        return f"xchg/{wid}/{self.r}"

    # ------------------------------------------------------------------ membership
    def connect(self, members: list[str], membership_epoch: int) -> None:
        """Map the regions of the current members (same fsdp_rank), close those of workers that left."""
        for w in list(self._open):
            if w not in members:
                self.lib.pb_ipc_close_handle(self._open.pop(w))
        hsize = self.lib.pb_ipc_handle_size()
        for w in members:
            if w == self.wid or w in self._open:
                continue
            info = json.loads(self.store.get(self._key(w)).decode())  # blocks until the peer has published (store timeout applies)
            if info["n"] != self.n:
                raise RuntimeError(f"elastic exchange: worker {w} shards {info['n']} elements, this worker {self.n} (different model / fsdp_size)")
            p = ctypes.c_void_p()
            _lib.check(self.lib.pb_ipc_open_handle(ctypes.create_string_buffer(bytes.fromhex(info["handle"]), hsize), ctypes.byref(p)),
                       f"pb_ipc_open_handle({w})")  # fmt: skip
            self._open[w] = int(p.value)
        self.members = list(members)
        self.index = self.members.index(self.wid)
        self._mem_epoch, self._seq = int(membership_epoch), 0

    def _bases(self) -> list[int]:
        return [self.base if w == self.wid else self._open[w] for w in self.members]

    def payload_ptrs(self) -> tuple[_lib.PeerPtrs, _lib.PeerPtrs]:
        bases = self._bases()
        return _lib.PeerPtrs.of(b + self.q_off for b in bases), _lib.PeerPtrs.of(b + self.s_off for b in bases)

    # ------------------------------------------------------------------ device-side sync
    def barrier(self, slot: int, stream: int) -> None:
        self._seq += 1
        epoch = ((self._mem_epoch << SEQ_BITS) + self._seq) & 0xFFFFFFFF
        flags = _lib.PeerPtrs.of(self._bases())
        _lib.check(self.lib.pb_barrier(ctypes.byref(flags), slot, self.index, epoch, self.err.data_ptr(), stream), "pb_barrier")

    def finish(self) -> None:
        """Synchronise and raise if a peer never showed up (the kernels left every optimizer tensor untouched in that case)."""
        code = int(self.err[0].item())
        if code != 0:
            self.err.zero_()
            raise RuntimeError(f"elastic exchange: a member of {self.members} did not reach the outer barrier (code {code})")

    def close(self) -> None:
        for p in self._open.values():
            self.lib.pb_ipc_close_handle(p)
        self._open.clear()
        self._whole = None
        if self.base:
            self.lib.pb_ipc_free(self.base)
            self.base = 0
