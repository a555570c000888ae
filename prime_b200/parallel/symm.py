"""Symmetric NVLink heap: one device allocation per rank, IPC-mapped into every peer of the box.

Every rank performs the same sequence of ``alloc`` calls, so a buffer lives at the same offset on
all ranks and ``peer_ptr(rank, tensor)`` is just ``peer_base[rank] + offset``.  Kernels in
``csrc/comm.cu`` take those peer pointers and load/store through NVSwitch directly — this is the
substrate for the fused reduce-scatter⊕AdamW⊕all-gather and int8 outer all-reduce⊕Nesterov paths.

Handle exchange uses whatever the caller provides (``torch.distributed`` object all-gather by
default, a TCPStore for the elastic multi-launch mode), so it needs no NCCL.
"""

from __future__ import annotations

import ctypes
from typing import Callable, Sequence

import torch

from ..ops import _lib


class _DevBuffer:
    """Expose a raw device pointer to torch through ``__cuda_array_interface__`` (zero copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,),
            "typestr": "|u1",
            "data": (ptr, False),
            "version": 3,
            "strides": None,
        }


class SymmetricHeap:
    FLAG_SLOTS = 4096  # uint32 flag words reserved at the start of the heap

    def __init__(
        self,
        nbytes: int,
        rank: int,
        world_size: int,
        exchange: Callable[[bytes], Sequence[bytes]],
        device: torch.device,
    ):
        self.lib = _lib.load()
        self.rank, self.world_size, self.device = rank, world_size, device
        self.nbytes = (nbytes + (1 << 21) - 1) & ~((1 << 21) - 1)
        base = ctypes.c_void_p()
        _lib.check(self.lib.pb_ipc_alloc(ctypes.byref(base), self.nbytes), "pb_ipc_alloc")
        self.base = int(base.value)
        hsize = self.lib.pb_ipc_handle_size()
        hbuf = ctypes.create_string_buffer(hsize)
        _lib.check(self.lib.pb_ipc_get_handle(self.base, hbuf), "pb_ipc_get_handle")
        handles = list(exchange(bytes(hbuf.raw)))
        assert len(handles) == world_size
        self.peer_base: list[int] = []
        for r, h in enumerate(handles):
            if r == rank:
                self.peer_base.append(self.base)
                continue
            p = ctypes.c_void_p()
            _lib.check(self.lib.pb_ipc_open_handle(ctypes.create_string_buffer(h, hsize), ctypes.byref(p)), "pb_ipc_open_handle")
            self.peer_base.append(int(p.value))
        # Rendezvous AFTER every rank has zeroed its own heap and mapped every peer: nobody may start writing into a peer's heap
        # while that peer's clearing memset (``pb_ipc_alloc``: asynchronous on the null stream, ≈10 ms for a 44 GB heap) is still
        # running, or while a peer is still inside cudaIpcOpenMemHandle on it. Seen with 44 GB heaps on 4 and 8 GPUs: bf16
        # weights of LATE layers (high addresses — the memset reaches them last), pushed by a faster peer, came out wrong on
        # some ranks before the first forward; 7 GB heaps never showed it. The synchronize makes the local memset complete,
        # the second exchange makes that true of every rank before anyone returns.
        torch.cuda.synchronize(device)
        if world_size > 1:
            exchange(b"mapped")
        self._whole = torch.as_tensor(_DevBuffer(self.base, self.nbytes), device=device)
        self._cursor = 0
        # control block: flags + error word + norm slots
        self.flags = self.alloc(self.FLAG_SLOTS, torch.int32)
        self.err = self.alloc(64, torch.int32)
        self.norm_slots = self.alloc(64, torch.float32)
        self._flag_cursor = 0

    # ------------------------------------------------------------------ allocation
    def alloc(self, numel: int, dtype: torch.dtype, align: int = 1024) -> torch.Tensor:
        esz = torch.empty((), dtype=dtype).element_size()
        start = (self._cursor + align - 1) // align * align
        end = start + numel * esz
        if end > self.nbytes:
            raise MemoryError(f"symmetric heap exhausted: need {end} of {self.nbytes} bytes")
        self._cursor = end
        return self._whole[start:end].view(dtype)

    def alloc_flags(self, n: int) -> int:
        """Reserve ``n`` consecutive flag slots; returns the first slot index."""
        s = self._flag_cursor
        if s + n > self.FLAG_SLOTS:
            raise MemoryError("out of flag slots")
        self._flag_cursor += n
        return s

    def offset_of(self, t: torch.Tensor) -> int:
        off = t.data_ptr() - self.base
        assert 0 <= off < self.nbytes, "tensor is not in the symmetric heap"
        return off

    def peer_ptr(self, rank: int, t: torch.Tensor) -> int:
        return self.peer_base[rank] + self.offset_of(t)

    def peers(self, ranks: Sequence[int], t: torch.Tensor) -> _lib.PeerPtrs:
        return _lib.PeerPtrs.of(self.peer_ptr(r, t) for r in ranks)

    # ------------------------------------------------------------------ sync helpers
    def barrier(self, ranks: Sequence[int], base_slot: int, epoch: int, stream: int) -> None:
        my_idx = list(ranks).index(self.rank)
        pp = self.peers(ranks, self.flags)
        _lib.check(
            self.lib.pb_barrier(ctypes.byref(pp), base_slot, my_idx, epoch & 0xFFFFFFFF, self.err.data_ptr(), stream),
            "pb_barrier",
        )

    def check_errors(self) -> None:
        """Raise if any device-side wait gave up (synchronises with the device). The comm kernels skip their stores when a wait
        fails, so the optimizer state is untouched but the step is void: callers must not continue or checkpoint past this."""
        code = int(self.err[0].item())
        if code != 0:
            why = {1: "timed out", 2: "was aborted by the host watchdog"}.get(code, f"failed (code {code})")
            raise RuntimeError(f"device-side peer wait {why} on rank {self.rank}: a rank of the NVLink group stalled or died")

    def clear_errors(self) -> None:
        self.err.zero_()

    def abort(self) -> None:
        """Host watchdog hook: make every device-side spin of this rank return immediately (a peer is known to be dead).
        Written from a side stream so it overtakes the spinning kernel."""
        if not hasattr(self, "_abort_stream"):
            self._abort_stream = torch.cuda.Stream(device=self.device, priority=-1)
            self._abort_word = torch.full((1,), 2, dtype=torch.int32, device=self.device)
        with torch.cuda.stream(self._abort_stream):
            self.err[:1].copy_(self._abort_word, non_blocking=True)

    def set_spin_timeout(self, seconds: float) -> None:
        """Bound of every device-side flag wait (default 20 s). Elastic jobs lower it to a heartbeat period."""
        _lib.check(self.lib.pb_set_spin_timeout_ms(int(max(1.0, seconds * 1e3))), "pb_set_spin_timeout_ms")

    def close(self) -> None:
        for r, p in enumerate(self.peer_base):
            if r != self.rank and p:
                self.lib.pb_ipc_close_handle(p)
        self.peer_base = []
        self._whole = None
        if self.base:
            self.lib.pb_ipc_free(self.base)
            self.base = 0


def dist_exchange(group=None) -> Callable[[bytes], Sequence[bytes]]:
    import torch.distributed as dist

    def _ex(mine: bytes) -> Sequence[bytes]:
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return [mine]
        out: list = [None] * dist.get_world_size(group)
        dist.all_gather_object(out, mine, group=group)
        return out

    return _ex
