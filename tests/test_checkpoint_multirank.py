"""CheckpointManager with several ranks (gloo, CPU): an asynchronously written checkpoint is published by poll() as soon as every shard
is on disk, and a failure on ANY rank — a writer thread, or rank 0's publish — raises on EVERY rank instead of leaving peers in a barrier."""

import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent(
    """
    import sys, time, torch
    import torch.distributed as dist
    sys.path.insert(0, {root!r})
    from prime_b200 import checkpoint as ck
    dist.init_process_group("gloo")
    rank, world, mode, root = dist.get_rank(), dist.get_world_size(), {mode!r}, {ckroot!r}
    mgr = ck.CheckpointManager(root, rank=rank, world_size=world, async_write=True)
    real_write, real_publish = ck.write_shard, mgr._publish
    if mode == "writer_fails" and rank == 1:
        def boom(*a, **k):
            raise OSError("disk full on rank 1")
        ck.write_shard = boom
    if mode == "publish_fails":
        def nope(step, meta):
            raise OSError("cannot rename")
        mgr._publish = nope  # only rank 0 ever calls it
    mgr.save(4, {{"w": torch.full((8,), float(rank))}}, {{"note": rank}}, {{"job": "t"}})
    outcome = "timeout"
    try:
        for _ in range(400):  # what the training loop does once per step
            if mgr.poll():
                outcome = "published"
                break
            time.sleep(0.01)
    except RuntimeError as e:
        outcome = "raised: " + str(e) + " <- " + repr(e.__cause__)
    open(f"{{root}}.outcome.{{rank}}", "w").write(f"OUTCOME {{rank}} {{outcome}}")  # per-rank file: two processes interleave on stdout
    if outcome == "published":
        assert (ck.step_dir(ck.Path(root), 4) / f"rank_{{rank:05d}}.pbck").is_file() and mgr._pending is None
        t, extra, meta = mgr.load(ck.step_dir(ck.Path(root), 4))
        assert float(t["w"][0]) == rank and extra["note"] == rank and meta["world_size"] == world
    mgr.wait()  # nothing in flight any more: must not block or raise
    dist.barrier()
    dist.destroy_process_group()
    """
)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(mode: str, tmp_path: Path) -> list[str]:
    script = tmp_path / f"worker_{mode}.py"
    script.write_text(WORKER.format(root=str(ROOT), mode=mode, ckroot=str(tmp_path / f"ck_{mode}")))
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]  # fmt: skip
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=180, env=env)  # a hang (the old behaviour) fails here
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return [Path(f"{tmp_path / ('ck_' + mode)}.outcome.{r}").read_text() for r in range(2)]


@pytest.mark.slow
def test_poll_publishes_when_every_shard_is_written(tmp_path):
    assert _run("ok", tmp_path) == ["OUTCOME 0 published", "OUTCOME 1 published"]
    d = tmp_path / "ck_ok"
    assert (d / "latest").read_text() == "step_000004" and not list(d.glob(".tmp-*"))


@pytest.mark.slow
def test_a_failed_writer_thread_raises_on_every_rank(tmp_path):
    out = _run("writer_fails", tmp_path)
    assert "failed on another rank" in out[0] and "failed on this rank" in out[1] and "disk full on rank 1" in out[1]
    assert not (tmp_path / "ck_writer_fails" / "latest").exists()


@pytest.mark.slow
def test_a_failed_publish_raises_on_every_rank_instead_of_hanging(tmp_path):
    out = _run("publish_fails", tmp_path)
    assert all("could not publish" in l for l in out) and "cannot rename" in out[0] and len(out) == 2
