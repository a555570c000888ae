"""Diagnostic: run a few steps of a large model in replicated-parameter mode and report, per step, loss / grad-norm and the first
bucket whose gradient, master or bf16 parameter buffer contains a non-finite value.

    python tools/debug_7b.py --model 7B --seq 2048 --micro-bs 1 --steps 4            # 1 GPU (F = 1, fused backend)
    torchrun --nproc-per-node 2 tools/debug_7b.py --model 7B --fsdp 2 ...
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from prime_b200.config import Config  # noqa: E402
from prime_b200.trainer import Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="7B")
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--micro-bs", type=int, default=1)
    ap.add_argument("--accum", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--fsdp", type=int, default=0)
    ap.add_argument("--reshard", type=int, default=0)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    F = a.fsdp or world
    cfg = Config.model_validate({
        "name_model": a.model, "data": {"seq_length": a.seq, "fake": True},
        "optim": {"batch_size": a.micro_bs * a.accum * F, "warmup_steps": 1000, "optim": {"lr": 3e-4}},
        "train": {"micro_bs": a.micro_bs, "reshard_after_forward": bool(a.reshard)}, "mesh": {"num_workers": world // F, "fsdp_size": F},
    })  # fmt: skip
    t = Trainer(cfg)
    eng = t.engine
    rank = t.mesh.world.rank
    rows = []

    def nonfinite(buf: torch.Tensor, chunk: int = 1 << 28) -> tuple[int, float]:
        n, mx = 0, 0.0
        for lo in range(0, buf.numel(), chunk):
            seg = buf[lo : lo + chunk]
            fin = torch.isfinite(seg)
            n += int((~fin).sum())
            mx = max(mx, float(torch.where(fin, seg, torch.zeros_like(seg)).abs().max()))
            del fin
        return n, mx

    # per-layer finiteness of the residual stream for every forward, and a checksum of the bf16 weights around the first step
    fw = []

    def hook(i):
        def f(mod, inp, out):
            h = out[0] if isinstance(out, tuple) else out
            if len(fw) < 4 * len(t.model.layers):
                fw.append((i, int((~torch.isfinite(h)).sum()), float(h.detach().float().abs().nan_to_num(0, 0, 0).max())))
        return f

    for i, layer in enumerate(t.model.layers):
        layer.register_forward_hook(hook(i))

    def wsum():
        return [float(p.detach().float().abs().sum()) for p in (t.model.tok_embeddings.weight, t.model.layers[0].attention.wqkv,
                                                                 t.model.layers[-1].feed_forward.w2, t.model.output)]

    torch.cuda.synchronize()
    print(json.dumps({"rank": rank, "weights_before": wsum(), "param_flat_nonfinite": nonfinite(eng.param_flat)[0]}), flush=True)
    for step in range(a.steps):
        r = t.inner_step()
        if step == 0:
            torch.cuda.synchronize()
            bad = [x for x in fw if x[1]]
            print(json.dumps({"rank": rank, "fwd_layers_seen": len(fw), "first_bad": bad[:2], "absmax_by_layer_first_fwd": [round(x[2], 2) for x in fw[: len(t.model.layers)]][::4],
                              "absmax_second_fwd": [round(x[2], 2) for x in fw[len(t.model.layers): 2 * len(t.model.layers)]][::4], "weights_after_step": wsum()}), flush=True)
        torch.cuda.synchronize()
        row = {"step": step + 1, "loss": float(r.loss), "gnorm": float(r.grad_norm)}
        if rank == 0:
            print(json.dumps(row), flush=True)
        bad = {}
        for name, buf, key in (("grad_flat", eng.grad_flat, "start"), ("param_flat", eng.param_flat, "pstart")):
            for b in eng.buckets:
                if b.kind != "flat" and name == "param_flat":
                    continue
                lo = getattr(b, key)
                n, mx = nonfinite(buf[lo : lo + b.size])
                if n:
                    bad.setdefault(name, []).append((b.name, n, mx))
                    break
        for name, buf in (("master", eng.master), ("exp_avg_sq", eng.exp_avg_sq), ("gshard", eng.gshard)):
            n, mx = nonfinite(buf)
            if n:
                bad[name] = (n, mx)
        row["nonfinite"] = bad
        row["rank"] = rank
        rows.append(row)
        if rank == 0 or bad:
            print(json.dumps(row), flush=True)
        if bad:
            break
    t.check_health()
    t.close()


if __name__ == "__main__":
    main()
