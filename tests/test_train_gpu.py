"""GPU: the train entrypoint end to end on the fused backend — async pinned-snapshot checkpoints and bit-exact resume."""

import json

import pytest
import torch

from prime_b200 import checkpoint as ck
from prime_b200.config import load_config
from prime_b200.train import train

pytestmark = pytest.mark.gpu

BASE = ["--name_model", "150M", "--data.seq_length", "256", "--optim.batch_size", "8", "--train.micro_bs", "4", "--optim.warmup_steps", "2",
        "--optim.total_steps", "8", "--diloco.inner_steps", "4", "--train.attn_impl", "native"]  # fmt: skip


def _losses(p):
    return {r["step"]: r["loss"] for r in map(json.loads, p.read_text().splitlines())}


def test_gpu_train_checkpoint_resume_exact(tmp_path):
    assert torch.cuda.is_available()
    a, b = tmp_path / "a", tmp_path / "b"
    out = train(load_config(BASE + ["--monitor.jsonl_path", str(a / "log.jsonl")]))
    assert out["step"] == 8 and out["loss"] < 11.0
    straight = _losses(a / "log.jsonl")
    assert straight[8] < straight[1]  # the synthetic stream is learnable
    train(load_config(BASE + ["--ckpt.path", str(b), "--ckpt.interval", "4", "--monitor.jsonl_path", str(b / "l1.jsonl")]), max_steps=4)
    assert ck.list_steps(b) == [4]
    shard, extra = ck.read_shard(b / "step_000004" / "rank_00000.pbck")
    assert {"master", "exp_avg", "exp_avg_sq", "theta0", "momentum", "rng_cpu", "rng_cuda"} == set(shard) and extra["trainer_step"] == 4
    train(load_config(BASE + ["--ckpt.path", str(b), "--ckpt.resume", "latest", "--monitor.jsonl_path", str(b / "l2.jsonl")]))
    resumed = {**_losses(b / "l1.jsonl"), **_losses(b / "l2.jsonl")}
    assert resumed == straight
