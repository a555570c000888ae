"""Forward-only reproduction of a NaN: the 7B model (same init as the trainer), the exact first micro-batches a given data rank
draws, per-layer finiteness of the residual stream, loss, then one backward with the gradient norm of the embedding input.

    python tools/debug_fwd.py --model 7B --seq 4096 --micro-bs 4 --data-rank 1 --data-world 4 --batches 2
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from prime_b200.data import FakeTokenDataset  # noqa: E402
from prime_b200.models.llama import build_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="7B")
    ap.add_argument("--seq", type=int, default=4096)
    ap.add_argument("--micro-bs", type=int, default=4)
    ap.add_argument("--data-rank", type=int, default=1)
    ap.add_argument("--data-world", type=int, default=4)
    ap.add_argument("--batches", type=int, default=2)
    ap.add_argument("--repeat", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    model = build_model(a.model, "llama2", device=dev, dtype=torch.bfloat16, seed=42, max_seq_len=max(a.seq, 128))
    ds = FakeTokenDataset(model.args.vocab_size, a.seq, 1337, a.data_rank, a.data_world)
    stats = {}

    def hook(name):
        def f(mod, inp, out):
            h = out[0] if isinstance(out, tuple) else out
            d = out[1] if isinstance(out, tuple) and out[1] is not None else None
            stats[name] = (int((~torch.isfinite(h)).sum()), float(h.float().abs().nan_to_num(0, 0, 0).max()),
                           None if d is None else int((~torch.isfinite(d)).sum()), None if d is None else float(d.float().abs().nan_to_num(0, 0, 0).max()))
        return f

    for i, layer in enumerate(model.layers):
        layer.register_forward_hook(hook(f"layer{i}"))
    for b in range(a.batches):
        x, y = ds.next_batch(a.micro_bs)
        tok = torch.from_numpy(x).to(dev)
        lab = torch.from_numpy(y).to(dev)
        for rep in range(a.repeat):
            stats.clear()
            loss = model.loss(tok, lab)
            torch.cuda.synchronize()
            bad = {k: v for k, v in stats.items() if v[0] or (v[2] or 0)}
            first = next(iter(bad.items()), None)
            print(json.dumps({"batch": b, "rep": rep, "loss": float(loss), "first_bad_layer": first, "n_bad_layers": len(bad),
                              "last_layer": stats.get(f"layer{len(model.layers) - 1}"), "tok_minmax": [int(tok.min()), int(tok.max())]}), flush=True)
            loss.backward()
            torch.cuda.synchronize()
            g = model.layers[0].attention.wqkv.grad
            print(json.dumps({"batch": b, "rep": rep, "wqkv0_grad_nonfinite": int((~torch.isfinite(g)).sum()), "wqkv0_grad_absmax": float(g.float().abs().nan_to_num(0, 0, 0).max())}), flush=True)
            model.zero_grad(set_to_none=True)


if __name__ == "__main__":
    main()
