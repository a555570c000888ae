"""Multi-process plumbing on CPU/gloo (BASELINE.json config 1: DiLoCo, 2 workers × 1 rank) and FSDP×DiLoCo meshes."""

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
    import json, os, sys, torch
    import torch.distributed as dist
    sys.path.insert(0, {root!r})
    from prime_b200.config import load_config
    from prime_b200.trainer import Trainer
    cfg = load_config({argv!r})
    t = Trainer(cfg)
    losses = []
    for i in range({steps}):
        r = t.inner_step()
        losses.append(float(r.loss))
    h = torch.tensor([t.engine.param_hash()], dtype=torch.float64)
    hs = [torch.zeros_like(h) for _ in range(dist.get_world_size())]
    dist.all_gather(hs, h)
    if dist.get_rank() == 0:
        print("RESULT " + json.dumps({{"losses": losses, "hashes": [float(x) for x in hs],
              "outer": t.outer.outer_step_count if t.outer else 0, "mesh": t.mesh.describe(),
              "wire": t.outer.last_bytes_on_wire if t.outer else 0}}))
    dist.barrier()
    dist.destroy_process_group()
    """
)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(nproc: int, argv: list[str], steps: int, tmp_path) -> dict:
    import json

    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT), argv=argv, steps=steps))
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]  # fmt: skip
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT ") :])


BASE = ["--name_model", "debugmodel", "--data.seq_length", "32", "--optim.batch_size", "4", "--train.micro_bs", "2",
        "--optim.warmup_steps", "2", "--optim.optim.lr", "3e-3"]  # fmt: skip


def test_diloco_two_workers_int8(tmp_path):
    out = _run(2, BASE + ["--diloco.inner_steps", "5", "--mesh.num_workers", "2"], 10, tmp_path)
    assert out["mesh"] == "dl2xfsdp1" and out["outer"] == 2
    # after an outer step all workers hold identical parameters
    assert out["hashes"][0] == pytest.approx(out["hashes"][1], rel=0, abs=0)
    assert out["losses"][-1] < out["losses"][0]
    assert out["wire"] > 0


def test_fsdp2_replicas_consistent(tmp_path):
    out = _run(2, BASE + ["--mesh.fsdp_size", "2"], 4, tmp_path)
    assert out["mesh"] == "dl1xfsdp2"
    assert out["hashes"][0] == pytest.approx(out["hashes"][1], rel=0, abs=0)


@pytest.mark.slow
def test_two_level_mesh_2x2(tmp_path):
    out = _run(4, BASE + ["--diloco.inner_steps", "3", "--mesh.num_workers", "2", "--mesh.fsdp_size", "2"], 6, tmp_path)
    assert out["mesh"] == "dl2xfsdp2" and out["outer"] == 2
    assert len(set(out["hashes"])) == 1
