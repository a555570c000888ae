"""Host-side pieces of the NVLS heap (prime_b200/parallel/multicast.py): descriptors travel between processes over Unix
sockets (SCM_RIGHTS) and still refer to the sender's open file; size rounding. The device path is in test_multigpu.py."""

import multiprocessing as mp
import os

import pytest

from prime_b200.parallel.multicast import FdExchange, nvls_available, round_up


def _worker(rank: int, world: int, q_in, q_out, tmp: str) -> None:
    def exchange(mine: bytes):  # a toy all-gather through the parent
        q_out.put((rank, mine))
        return q_in.get()

    fdx = FdExchange(rank, world, exchange, timeout_s=20)
    try:
        path = os.path.join(tmp, f"payload-{rank}")
        with open(path, "w") as f:
            f.write(f"hello from {rank}")
        fd = os.open(path, os.O_RDONLY)
        extra = os.open(path, os.O_RDONLY) if rank == 0 else None  # rank 0 shares two descriptors, like (memory, multicast)
        got = fdx.all_gather([fd, extra] if extra is not None else [fd])
        texts = {}
        for r, fds in enumerate(got):
            assert len(fds) == (2 if r == 0 else 1)
            texts[r] = os.pread(fds[0], 64, 0).decode()  # pread: receivers share ONE open file description (and its offset)
            if r != rank:
                for x in fds:
                    os.close(x)
        os.close(fd)
        # second round on the same sockets (the heap does memory first, the multicast object second); only rank 0 contributes
        again = fdx.all_gather([os.open(path, os.O_RDONLY)] if rank == 0 else [])
        assert [len(x) for x in again] == [1] + [0] * (world - 1)
        q_out.put((rank, texts))
    finally:
        fdx.close()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("world", [2, 4])
def test_descriptors_cross_process_boundaries(tmp_path, world):
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    q_ins = [ctx.Queue() for _ in range(world)]
    procs = [ctx.Process(target=_worker, args=(r, world, q_ins[r], q_out, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    try:
        paths = dict(q_out.get(timeout=60) for _ in range(world))  # socket paths, one per rank
        for q in q_ins:
            q.put([paths[r] for r in range(world)])
        results = dict(q_out.get(timeout=60) for _ in range(world))
    finally:
        for p in procs:
            p.join(30)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    want = {r: f"hello from {r}" for r in range(world)}
    assert all(results[r] == want for r in range(world))  # every rank read every rank's file through a received descriptor


def test_rounding_and_probe():
    assert round_up(1, 2 << 20) == 2 << 20 and round_up(2 << 20, 2 << 20) == 2 << 20 and round_up((2 << 20) + 1, 2 << 20) == 4 << 20
    assert nvls_available() in (True, False)  # never raises, whatever the box
