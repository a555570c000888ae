"""Offline evaluation of a trained model: perplexity over a tokenised corpus and (small-scale) sampling.

    python -m prime_b200.eval ppl --ckpt runs/x/ckpt/step_001000 --model 1B --data data/c4_val --seq 1024 --batches 50
    python -m prime_b200.eval ppl --hf path/to/llama --data data/c4_val
    python -m prime_b200.eval generate --ckpt … --model 1B --prompt-ids 1,306,626 --max-new 32

The model is rebuilt on one device from the per-rank optimizer shards (``checkpoint.assemble_full_model``) or from a Hugging
Face directory (``models.hf.load_hf_dir``); the forward runs through the same ops as training (native kernels on a B200, the
reference ops on the CPU).  Generation recomputes the whole prefix for every token — there is no KV cache in a training engine —
and pads the sequence to a multiple of 128 on CUDA (causal attention makes right padding invisible to the position that is read),
which is fine for inspecting a checkpoint and not a serving path.
"""

from __future__ import annotations

import argparse
import json
import math

import torch
import torch.nn.functional as F

from .data import MemmapTokenDataset
from .models.llama import Transformer


@torch.no_grad()
def perplexity(model: Transformer, dataset, *, batches: int, batch_size: int, device: torch.device | str) -> dict:
    """Mean next-token loss (nats) and perplexity over ``batches`` × ``batch_size`` windows of ``dataset``."""
    model.eval()
    tot, n = 0.0, 0
    for _ in range(batches):
        x, y = dataset.next_batch(batch_size)
        tok = torch.from_numpy(x).to(device)
        tgt = torch.from_numpy(y).to(device)
        logits = model(tok)
        tot += float(F.cross_entropy(logits.float().reshape(-1, logits.shape[-1]), tgt.reshape(-1), reduction="sum"))
        n += tgt.numel()
    loss = tot / max(n, 1)
    return {"loss": loss, "perplexity": math.exp(min(loss, 50.0)), "tokens": n}


@torch.no_grad()
def generate(model: Transformer, prompt: list[int], *, max_new: int, temperature: float = 0.0, top_k: int = 0, seed: int = 0,
             eos_id: int | None = None) -> list[int]:  # fmt: skip
    """Greedy (temperature 0) or top-k / temperature sampling; returns prompt + continuation."""
    model.eval()
    dev = next(model.parameters()).device
    ids = list(prompt)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    pad_to = 128 if dev.type == "cuda" else 1
    for _ in range(max_new):
        n = len(ids)
        if n >= model.args.max_seq_len:
            break
        padded = ids + [0] * ((-n) % pad_to)
        logits = model(torch.tensor([padded], dtype=torch.int64, device=dev))[0, n - 1].float().cpu()
        if temperature <= 0:
            nxt = int(logits.argmax())
        else:
            logits = logits / temperature
            if top_k > 0:
                kth = torch.topk(logits, min(top_k, logits.numel())).values[-1]
                logits = logits.masked_fill(logits < kth, float("-inf"))
            nxt = int(torch.multinomial(torch.softmax(logits, -1), 1, generator=gen))
        ids.append(nxt)
        if eos_id is not None and nxt == eos_id:
            break
    return ids


def load_model(a, device: str) -> Transformer:
    dtype = torch.bfloat16 if device.startswith("cuda") else torch.float32
    if a.hf:
        from .models.hf import load_hf_dir

        return load_hf_dir(a.hf, device=device, dtype=dtype)
    from .checkpoint import assemble_full_model

    return assemble_full_model(a.ckpt, a.model, a.type_model).to(device=device, dtype=dtype)


def main(argv: list[str] | None = None) -> dict:
    ap = argparse.ArgumentParser(prog="python -m prime_b200.eval", description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    for name in ("ppl", "generate"):
        p = sub.add_parser(name)
        src = p.add_mutually_exclusive_group(required=True)
        src.add_argument("--ckpt", help="checkpoint step directory written by the trainer")
        src.add_argument("--hf", help="Hugging Face Llama checkpoint directory")
        p.add_argument("--model", default="1B")
        p.add_argument("--type-model", default="llama2", choices=["llama2", "llama3"])
        p.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    p = sub.choices["ppl"]
    p.add_argument("--data", required=True, help="token files / directory / glob (see tools/tokenize_corpus.py)")
    p.add_argument("--seq", type=int, default=1024)
    p.add_argument("--batches", type=int, default=20)
    p.add_argument("--batch-size", type=int, default=8)
    g = sub.choices["generate"]
    g.add_argument("--prompt-ids", required=True, help="comma-separated token ids")
    g.add_argument("--max-new", type=int, default=32)
    g.add_argument("--temperature", type=float, default=0.0)
    g.add_argument("--top-k", type=int, default=0)
    g.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)
    model = load_model(a, a.device)
    if a.cmd == "ppl":
        ds = MemmapTokenDataset(a.data, a.seq, vocab_size=model.args.vocab_size, shuffle=False)
        out = perplexity(model, ds, batches=min(a.batches, max(1, ds.windows // a.batch_size)), batch_size=a.batch_size, device=a.device)
    else:
        ids = generate(model, [int(x) for x in a.prompt_ids.split(",")], max_new=a.max_new, temperature=a.temperature, top_k=a.top_k, seed=a.seed)
        out = {"ids": ids}
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    main()
