"""NVLS heap: the symmetric heap re-built on CUDA virtual-memory management so that NVSwitch can multicast into it and
reduce out of it (``multimem.st`` / ``multimem.ld_reduce``; kernels and driver plumbing in ``csrc/multicast.cu``).

What changes against ``symm.SymmetricHeap`` (cudaIpc): the allocation is a VMM physical handle exported as a POSIX file
descriptor; peers import it and map it (unicast peer pointers, exactly as before), and additionally every rank binds its
allocation to ONE multicast object and maps that — a second address range (``mc_base``) where a store reaches all copies and
a ``ld_reduce`` returns their sum. File descriptors cannot travel through ``torch.distributed``; they go through Unix sockets
with ``SCM_RIGHTS`` (``FdExchange``) — rendezvous of the socket paths uses the same ``exchange`` callable as the IPC heap.

Opt-in (``PB_NVLS=1`` in the engine, or construct it directly): the switch's summation order differs from the rank-ordered
peer loads, so results are not bitwise identical to the default path. First measurements belong to round 2; the CPU-testable
parts (fd exchange, layout arithmetic) are covered by ``tests/test_multicast_cpu.py``, the device path by
``tests/test_multigpu.py`` (skipped where multicast is unsupported).

No counterpart in the reference (SURVEY §5 "distributed communication backend": absent); the design follows SURVEY §7.1 step 5.
"""

from __future__ import annotations

import ctypes
import os
import socket
import struct
import tempfile
from typing import Callable, Sequence

import torch

from ..ops import _lib
from .symm import SymmetricHeap, _DevBuffer

_HELLO = struct.Struct("!III")  # (round, sender rank, number of fds)


class FdExchange:
    """All-to-all of open file descriptors between the processes of one box.

    Every rank listens on a Unix socket; ``all_gather(fds)`` sends this rank's descriptors to every peer and returns, for
    each rank, the descriptors received from it (its own are returned as given). Receivers get NEW descriptor numbers that
    refer to the same open file descriptions — that is what ``cuMemImportFromShareableHandle`` needs.
    """

    def __init__(self, rank: int, world_size: int, exchange: Callable[[bytes], Sequence[bytes]], timeout_s: float = 60.0):
        self.rank, self.world_size, self.timeout_s = rank, world_size, timeout_s
        self._dir = tempfile.mkdtemp(prefix="pb-fd-")
        self.path = os.path.join(self._dir, f"r{rank}.sock")
        self._srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self._srv.bind(self.path)
        os.chmod(self.path, 0o600)
        self._srv.listen(world_size)
        self._srv.settimeout(timeout_s)
        self.paths = [p.decode() for p in exchange(self.path.encode())]  # also a barrier: every socket is listening now
        assert len(self.paths) == world_size
        self._round = 0
        self._early: dict[tuple[int, int], list[int]] = {}  # (round, sender) → descriptors of a peer that is a round ahead

    def all_gather(self, fds: Sequence[int]) -> list[list[int]]:
        got: list[list[int] | None] = [None] * self.world_size
        got[self.rank] = list(fds)
        self._round += 1
        rnd = self._round
        # phase 1 — post to every peer. connect() completes against the listen backlog and the message fits the socket buffer,
        # so nobody blocks here even though no one is accepting yet (waiting for an ack at this point would deadlock).
        outgoing = []
        try:
            for r, path in enumerate(self.paths):
                if r == self.rank:
                    continue
                c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                outgoing.append(c)
                c.settimeout(self.timeout_s)
                c.connect(path)
                socket.send_fds(c, [_HELLO.pack(rnd, self.rank, len(fds))], list(fds))
            # phase 2 — collect from every peer, acknowledging each. A peer that already finished this round may be posting the
            # next one: its message is kept for then (rounds never skip, so one round of look-ahead is all there can be).
            for r in range(self.world_size):
                if (rnd, r) in self._early:
                    got[r] = self._early.pop((rnd, r))
            while any(g is None for g in got):
                conn, _addr = self._srv.accept()
                with conn:
                    conn.settimeout(self.timeout_s)
                    msg, rfds, _flags, _a = socket.recv_fds(conn, _HELLO.size, 64)
                    their_round, sender, n = _HELLO.unpack(msg)
                    if len(rfds) != n:
                        raise RuntimeError(f"rank {sender} announced {n} descriptors, {len(rfds)} arrived")
                    if their_round == rnd:
                        got[sender] = list(rfds)
                    else:
                        self._early[(their_round, sender)] = list(rfds)
                    conn.send(b"\x01")
            # phase 3 — our descriptors have arrived everywhere once every peer has acknowledged: the caller may close them
            for c in outgoing:
                if c.recv(1) != b"\x01":
                    raise RuntimeError("a peer closed the descriptor channel without acknowledging")
        finally:
            for c in outgoing:
                c.close()
        return [g if g is not None else [] for g in got]

    def close(self) -> None:
        self._srv.close()
        try:
            os.unlink(self.path)
            os.rmdir(self._dir)
        except OSError:
            pass


def nvls_available(device_index: int = 0) -> bool:
    """True when the native library is loaded on a GPU box whose device can join a multicast object."""
    if not torch.cuda.is_available():
        return False
    try:
        return _lib.load().pb_mc_supported(device_index) == 1
    except Exception:
        return False


class MulticastHeap(SymmetricHeap):
    """``SymmetricHeap`` with the same allocation API, plus ``mc_ptr(t)``: the multicast address of a heap tensor."""

    def __init__(self, nbytes: int, rank: int, world_size: int, exchange: Callable[[bytes], Sequence[bytes]], device: torch.device):
        lib = self.lib = _lib.load()
        self.rank, self.world_size, self.device = rank, world_size, device
        dev = device.index if device.index is not None else torch.cuda.current_device()
        self._dev = dev
        if lib.pb_mc_supported(dev) != 1:
            raise RuntimeError("this device cannot join an NVSwitch multicast object (pb_mc_supported != 1)")
        gran = ctypes.c_size_t()
        _lib.check(lib.pb_vmm_granularity(dev, world_size, nbytes, ctypes.byref(gran)), "pb_vmm_granularity")
        self.granularity = int(gran.value)
        self.nbytes = round_up(nbytes, self.granularity)

        fdx = FdExchange(rank, world_size, exchange)
        try:
            # 1. physical memory, shared by descriptor; every peer's copy mapped for unicast access
            mem, fd = ctypes.c_uint64(), ctypes.c_int()
            _lib.check(lib.pb_vmm_create(dev, self.nbytes, ctypes.byref(mem), ctypes.byref(fd)), "pb_vmm_create")
            self._mem = int(mem.value)
            self._peer_handles: list[int] = []
            self.peer_base = []
            for r, fds in enumerate(fdx.all_gather([fd.value])):
                if r == rank:
                    handle = self._mem
                else:
                    h = ctypes.c_uint64()
                    _lib.check(lib.pb_vmm_import(fds[0], ctypes.byref(h)), "pb_vmm_import")
                    os.close(fds[0])
                    handle = int(h.value)
                    self._peer_handles.append(handle)
                p = ctypes.c_void_p()
                _lib.check(lib.pb_vmm_map(handle, self.nbytes, self.granularity, dev, ctypes.byref(p)), "pb_vmm_map")
                self.peer_base.append(int(p.value))
            os.close(fd.value)
            self.base = self.peer_base[rank]

            # 2. one multicast object for the box: rank 0 creates, everyone adds its device, then (barrier) binds its memory
            mc, mfd = ctypes.c_uint64(), ctypes.c_int(-1)
            if rank == 0:
                _lib.check(lib.pb_mc_create(world_size, self.nbytes, ctypes.byref(mc), ctypes.byref(mfd)), "pb_mc_create")
            shared = fdx.all_gather([mfd.value] if rank == 0 else [])
            if rank == 0:
                os.close(mfd.value)
            else:
                _lib.check(lib.pb_vmm_import(shared[0][0], ctypes.byref(mc)), "pb_vmm_import(multicast)")
                os.close(shared[0][0])
            self._mc = int(mc.value)
            _lib.check(lib.pb_mc_add_device(self._mc, dev), "pb_mc_add_device")
            exchange(b"added")  # cuMulticastBindMem requires that ALL devices have been added
            _lib.check(lib.pb_mc_bind(self._mc, self._mem, self.nbytes), "pb_mc_bind")
            p = ctypes.c_void_p()
            _lib.check(lib.pb_vmm_map(self._mc, self.nbytes, self.granularity, dev, ctypes.byref(p)), "pb_vmm_map(multicast)")
            self.mc_base = int(p.value)
            exchange(b"bound")
        finally:
            fdx.close()

        self._whole = torch.as_tensor(_DevBuffer(self.base, self.nbytes), device=device)
        self._whole.zero_()
        self._cursor = 0
        self.flags = self.alloc(self.FLAG_SLOTS, torch.int32)
        self.err = self.alloc(64, torch.int32)
        self.norm_slots = self.alloc(64, torch.float32)
        self._flag_cursor = 0
        # flag words of the multimem barriers (one per block of the all-reduce grid) and the launch counter they key on
        self.mc_flags = self.alloc(max(256, lib.pb_mc_all_reduce_grid()), torch.int32)
        self._mc_epoch = 0
        torch.cuda.synchronize(device)
        exchange(b"zeroed")  # nobody signals into a heap that is still being cleared

    # ------------------------------------------------------------------ addresses
    def mc_ptr(self, t: torch.Tensor) -> int:
        return self.mc_base + self.offset_of(t)

    # ------------------------------------------------------------------ collectives
    def all_reduce_(self, t: torch.Tensor, stream: int | None = None) -> torch.Tensor:
        """In-place sum over all ranks of a heap tensor (f32 or bf16; bytes a multiple of 16). Every rank must call it with a
        tensor at the same heap offset, in the same order — the barrier counters advance once per call."""
        if t.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"NVLS all-reduce supports float32 and bfloat16, not {t.dtype}")
        nbytes = t.numel() * t.element_size()
        if nbytes % 16 or self.offset_of(t) % 16:
            raise ValueError("NVLS all-reduce needs a 16-byte aligned tensor whose size is a multiple of 16 bytes")
        self._mc_epoch += 1
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.pb_mc_all_reduce(self.mc_ptr(t), nbytes, 1 if t.dtype == torch.bfloat16 else 0, self.mc_ptr(self.mc_flags),
                                             self.mc_flags.data_ptr(), self.rank, self.world_size, self._mc_epoch, self.err.data_ptr(), s),
                   "pb_mc_all_reduce")  # fmt: skip
        return t

    # ------------------------------------------------------------------ teardown
    def close(self) -> None:
        if not getattr(self, "base", 0):
            return
        lib = self.lib
        self._whole = None
        lib.pb_vmm_unmap(self.mc_base, self.nbytes)
        lib.pb_mc_unbind(self._mc, self._dev, self.nbytes)
        for p in self.peer_base:
            lib.pb_vmm_unmap(p, self.nbytes)
        for h in self._peer_handles:
            lib.pb_vmm_release(h)
        lib.pb_vmm_release(self._mc)
        lib.pb_vmm_release(self._mem)
        self.peer_base, self.base, self.mc_base = [], 0, 0


def round_up(n: int, multiple: int) -> int:
    return (n + multiple - 1) // multiple * multiple
