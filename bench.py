#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): tokens/sec, Llama-1B DiLoCo H=100 on 1/2/4/8 B200.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29501 bench.py --gpus 8 --steps 10 --warmup 3

Mesh: N=1 → 1 worker × 1 GPU;  N>=2 → (N/2) DiLoCo workers × 2-GPU FSDP (config 2 of BASELINE.json).
Weak scaling: every GPU processes the same number of tokens per step for every N.

A "step" is one full inner optimizer step (ACCUM micro-batches fwd+bwd, gradient reduce-scatter,
clip, AdamW, parameter all-gather).  The outer DiLoCo step (int8 pseudo-gradient all-reduce ⊕
Nesterov) runs every H=100 inner steps *and* is forced at the last step of each timed region so its
cost is always inside the measurement (pessimistic when K < H).

Two timed regions of K steps each, both through the public API ``Trainer.inner_step()`` and both
including the per-micro-batch pinned host→device input copies:
  * device-timed (CUDA events, max over ranks)                          → ``value``
  * end-to-end: same, plus a device→host read of the loss every step    → ``e2e``
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

MODEL = "1B"
SEQ = 1024
MICRO_BS = 16
ACCUM = 4
H = 100


def reference_arm() -> None:
    """The mounted reference is the Prime CLI/SDK: there is no trainer in it to time."""
    ref = ROOT / "baseline" / "_ref"
    why = (
        "reference (PrimeIntellect-ai/prime @ d892ed8) is a pure-Python HTTP CLI/SDK with no model, trainer, "
        "GPU code or tokens/s benchmark (0 files import torch); nothing in it can run this metric"
    )
    if not ref.exists():
        why = "baseline/_ref not installed; and " + why
    print(json.dumps({"impl": "reference", "unavailable": why}))


from prime_b200.utils.clocks import ClockSampler  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--micro-bs", type=int, default=MICRO_BS)
    ap.add_argument("--accum", type=int, default=ACCUM)
    ap.add_argument("--seq", type=int, default=SEQ)
    ap.add_argument("--no-fused-comm", action="store_true", help="baseline B0: NCCL collectives instead of fused P2P kernels")
    ap.add_argument("--attn", default="auto")
    ap.add_argument("--fp8", action="store_true", help="NON-headline: MXFP8 forward/dgrad GEMMs (reported with dtype 'mxfp8+bf16')")
    ap.add_argument("--graphs", type=int, default=0, help="capture each micro-step (fwd+bwd) in a CUDA graph")
    args = ap.parse_args()

    if args.impl == "reference":
        reference_arm()
        return

    import torch
    import torch.distributed as dist

    from prime_b200 import ops
    from prime_b200.config import Config
    from prime_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    warmup = max(3, args.warmup)
    K = max(1, args.steps)

    fsdp = 1 if world == 1 else 2
    workers = world // fsdp
    cfg = Config.model_validate(
        {
            "name_model": args.model,
            "type_model": "llama2",
            "data": {"seq_length": args.seq, "fake": True},
            "optim": {"batch_size": args.micro_bs * args.accum * fsdp, "warmup_steps": 10, "total_steps": 100000,
                      "optim": {"lr": 4e-4}},
            "train": {"micro_bs": args.micro_bs, "fused_comm": not args.no_fused_comm, "attn_impl": args.attn,
                      "cuda_graphs": bool(args.graphs), "fp8": args.fp8},
            "diloco": {"inner_steps": H, "compression": "int8", "outer_lr": 0.7},
            "mesh": {"num_workers": workers, "fsdp_size": fsdp},
        }
    )  # fmt: skip
    trainer = Trainer(cfg)
    dev = trainer.device
    rank = trainer.mesh.world.rank

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_region(n_steps: int, read_loss: bool) -> tuple[float, float, int]:
        """Returns (device ms, host s, launches) for n_steps inner steps incl. >= 1 outer step."""
        sync_all()
        ops.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        did_outer = False
        last = None
        for i in range(n_steps):
            r = trainer.inner_step()
            did_outer |= r.did_outer
            if i == n_steps - 1 and not did_outer and trainer.outer is not None:
                trainer.outer.step()
            if read_loss:
                last = float(r.loss.item())  # device→host read of the step result, every step
        e1.record()
        sync_all()
        host = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        hs = torch.tensor([host], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(hs, op=dist.ReduceOp.MAX)
        if read_loss and last is not None and not (last == last):
            raise RuntimeError("loss is NaN")
        return float(ms.item()), float(hs.item()), ops.launch_count()

    for _ in range(warmup):
        trainer.inner_step()
    if trainer.outer is not None:
        trainer.outer.step()  # warm the outer path too
    sync_all()

    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    dev_ms, _, launches = run_region(K, read_loss=False)
    e2e_ms, e2e_host_s, _ = run_region(K, read_loss=True)
    clocks = sampler.finish() if rank == 0 else {}

    tokens = trainer.tokens_per_step * K
    value = tokens / (dev_ms / 1e3)
    e2e_value = tokens / max(e2e_host_s, e2e_ms / 1e3)
    h2d = trainer.loader.h2d_bytes_per_batch * trainer.accum
    if rank == 0:
        flops = trainer.flops_per_step() * K / world
        line = {
            "metric": "tokens/sec Llama-1B DiLoCo H=100",
            "value": round(value, 1),
            "unit": "tokens/s",
            "n_gpus": world,
            "steps": K,
            "warmup": warmup,
            "ms_per_step": round(dev_ms / K, 3),
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "bf16" if not args.fp8 else "mxfp8+bf16 (non-headline)",
            "data": "synthetic tokens, random-init weights",
            "impl": "ours",
            "config": {
                "model": f"Llama-{args.model} (dim 2048, 18 layers, 16 heads, vocab 32000)" if args.model == "1B" else args.model,
                "global_batch": args.micro_bs * args.accum * world,
                "seq_len": args.seq,
                "micro_bs": args.micro_bs,
                "grad_accum": args.accum,
                "parallelism": f"diloco{workers}xfsdp{fsdp}",
                "diloco_H": H,
                "outer": "int8 all-gather + Nesterov, forced >=1 per timed region",
                "comm": "fused P2P kernels" if not args.no_fused_comm else "NCCL collectives (B0)",
                "cuda_graphs": bool(args.graphs),
                "l2": "working set (params+activations, >20 GB/step) far larger than the 126 MB L2; no flush needed",
            },
            "clocks": clocks,
            "e2e": {
                "value": round(e2e_value, 1),
                "unit": "tokens/s",
                "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4,
                "ms_per_step": round(max(e2e_host_s * 1e3, e2e_ms) / K, 3),
            },
            "gpu_launches": launches,
            "mfu_of_measured_sustained_peak": None,
        }
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            line["mfu_of_measured_sustained_peak"] = round(flops / (dev_ms / 1e3) / (peaks["bf16_tflops_sustained"] * 1e12), 4)
        except Exception:
            pass
        print(json.dumps(line), flush=True)
    trainer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
