#!/usr/bin/env python
"""Loss curve of the native engine against the stock-PyTorch B0 arm: same architecture, SAME initial weights (copied), same
synthetic token stream, same AdamW hyper-parameters and schedule (linear warm-up, then constant), same batch plan.

    python tools/loss_curve.py --model 150M --steps 300 --out profiles/loss_curve_150M_r2.json

Single GPU (fused engine with F = 1 vs FSDP2 on a 1-rank mesh). The two curves must track each other within bf16 noise; the
JSON keeps both series, their difference statistics, and the per-step time of both arms. VERDICT r1 #7."""

from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="150M")
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--micro-bs", type=int, default=8)
    ap.add_argument("--accum", type=int, default=2)
    ap.add_argument("--lr", type=float, default=4e-4)
    ap.add_argument("--out", default="gpurun_out/loss_curve.json")
    a = ap.parse_args()

    from baseline.torch_b0 import B0Config, B0Trainer
    from prime_b200.config import Config
    from prime_b200.trainer import Trainer
    from prime_b200.utils.clocks import ClockSampler

    sampler = ClockSampler(0)
    sampler.start()
    cfg = Config.model_validate({
        "name_model": a.model, "seed": 42, "data": {"seq_length": a.seq, "fake": True, "seed": 42},
        "optim": {"batch_size": a.micro_bs * a.accum, "warmup_steps": 10, "total_steps": 10**6, "sched_type": "constant", "optim": {"lr": a.lr}},
        "train": {"micro_bs": a.micro_bs},
    })  # fmt: skip
    tr = Trainer(cfg)
    # B0 with the SAME initial weights: build it, then copy the native model's tensors into its (1-rank FSDP2) parameters
    b0 = B0Trainer(B0Config(model=a.model, seq=a.seq, micro_bs=a.micro_bs, accum=a.accum, workers=1, fsdp=1, lr=a.lr, seed=42, diloco=False))
    loc = b0._local
    with torch.no_grad():
        m, r = tr.model, b0.model
        loc(r.tok_embeddings.weight).copy_(m.tok_embeddings.weight.float())
        for x, y in zip(m.layers, r.layers):
            loc(y.wqkv.weight).copy_(x.attention.wqkv.float())
            loc(y.wo.weight).copy_(x.attention.wo.float())
            loc(y.w13.weight).copy_(x.feed_forward.w13.float())
            loc(y.w2.weight).copy_(x.feed_forward.w2.float())
            loc(y.attention_norm.weight).copy_(x.attention_norm.weight.float())
            loc(y.ffn_norm.weight).copy_(x.ffn_norm.weight.float())
        loc(r.norm.weight).copy_(m.norm.weight.float())
        loc(r.output.weight).copy_(m.output.float())
    ours, base = [], []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ours.append(tr.inner_step().loss)
    ours = [float(x) for x in torch.stack(ours).cpu()]
    torch.cuda.synchronize()
    t_ours = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        base.append(b0.inner_step())
    base = [float(x) for x in torch.stack(base).cpu()]
    torch.cuda.synchronize()
    t_b0 = time.perf_counter() - t0
    d = [x - y for x, y in zip(ours, base)]
    tail = slice(a.steps // 2, None)
    out = {
        "model": a.model, "steps": a.steps, "seq": a.seq, "micro_bs": a.micro_bs, "accum": a.accum, "lr": a.lr,
        "note": "same init (copied), same token stream (seed 42), AdamW(0.9, 0.95, wd 0.1, clip 1.0), 10-step linear warm-up then constant",
        "first_loss": {"ours": ours[0], "b0": base[0]},
        "last_loss": {"ours": ours[-1], "b0": base[-1]},
        "mean_abs_diff": sum(abs(x) for x in d) / len(d),
        "max_abs_diff": max(abs(x) for x in d),
        "mean_diff_second_half": sum(d[tail]) / len(d[tail]),
        "mean_loss_second_half": {"ours": sum(ours[tail]) / len(ours[tail]), "b0": sum(base[tail]) / len(base[tail])},
        "s_per_step": {"ours": t_ours / a.steps, "b0": t_b0 / a.steps},
        "clocks": sampler.finish(),
        "ours": [round(x, 5) for x in ours],
        "b0": [round(x, 5) for x in base],
    }  # fmt: skip
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(out, indent=1))
    print(json.dumps({k: v for k, v in out.items() if k not in ("ours", "b0")}, indent=1))
    tr.close()


if __name__ == "__main__":
    main()
