"""2-GPU elastic run: two independently launched single-GPU workers (own process worlds, fused intra-worker engine) meet through
the global store, exchange int8 pseudo-gradients over a per-epoch NCCL group, and end with identical parameters."""

import json
import os
import subprocess
import sys
import time
from pathlib import Path

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = Path(__file__).resolve().parent.parent


def test_two_elastic_workers_on_two_gpus(tmp_path):
    from prime_b200.parallel import elastic as el

    served = el.serve(0)
    argv = ["--name_model", "150M", "--data.seq_length", "256", "--optim.batch_size", "4", "--train.micro_bs", "4", "--optim.warmup_steps", "2",
            "--optim.total_steps", "8", "--diloco.inner_steps", "4", "--mesh.elastic", "true", "--mesh.num_workers", "2",
            "--mesh.heartbeat_interval_s", "0.2", "--mesh.heartbeat_timeout_s", "5", "--train.log_model_hash", "true", "--train.attn_impl", "native"]  # fmt: skip
    procs = {}
    for i, name in enumerate(("w0", "w1")):
        env = {**os.environ, "PYTHONPATH": str(ROOT), "CUDA_VISIBLE_DEVICES": str(i), "GLOBAL_ADDR": "127.0.0.1", "GLOBAL_PORT": str(served.port),
               "GLOBAL_UNIQUE_ID": name, "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(29700 + i)}  # fmt: skip
        log = open(tmp_path / f"{name}.log", "w")
        procs[name] = subprocess.Popen([sys.executable, "-m", "prime_b200.train", *argv, "--monitor.jsonl_path", str(tmp_path / f"{name}.jsonl")],
                                       cwd=ROOT, env=env, stdout=log, stderr=subprocess.STDOUT)  # fmt: skip
    try:
        deadline = time.time() + 300
        for name, p in procs.items():
            rc = p.wait(timeout=max(1, deadline - time.time()))
            assert rc == 0, (tmp_path / f"{name}.log").read_text()[-3000:]
    finally:
        for p in procs.values():
            if p.poll() is None:
                p.kill()
    rows = {n: [json.loads(x) for x in (tmp_path / f"{n}.jsonl").read_text().splitlines()] for n in procs}
    last = {n: [r for r in rows[n] if r.get("outer")][-1] for n in rows}
    assert last["w0"]["step"] == last["w1"]["step"] == 8 and last["w0"]["workers"] == 2
    assert last["w0"]["outer_bytes"] > 0 and last["w0"]["param_hash"] == last["w1"]["param_hash"]
    del served.store
