"""Data path: synthetic and memory-mapped token datasets, rank striding, the prefetcher's consumer-side checkpoint position."""

import numpy as np
import pytest
import torch

from prime_b200.data import FakeTokenDataset, MemmapTokenDataset, PinnedPrefetcher


def test_fake_dataset_is_deterministic_and_learnable():
    a, b = FakeTokenDataset(1000, 16, seed=3, rank=0), FakeTokenDataset(1000, 16, seed=3, rank=0)
    xa, ya = a.next_batch(4)
    xb, yb = b.next_batch(4)
    assert np.array_equal(xa, xb) and np.array_equal(ya, yb) and xa.shape == (4, 16)
    assert np.array_equal(xa[:, 1:], ya[:, :-1])  # labels are the inputs shifted by one
    assert ((ya - xa) % 1000).max() <= 3 and ((ya - xa) % 1000).min() >= 1  # next token = previous + {1,2,3}: learnable
    other = FakeTokenDataset(1000, 16, seed=3, rank=1).next_batch(4)[0]
    assert not np.array_equal(xa, other)  # ranks draw different streams


def test_memmap_dataset_windows_and_rank_striding(tmp_path):
    toks = np.arange(0, 8 * 10 + 1, dtype=np.uint16)  # 8 windows of seq 10
    f = tmp_path / "a.bin"
    toks.tofile(f)
    d0, d1 = MemmapTokenDataset(str(f), 10, rank=0, world=2, shuffle=False), MemmapTokenDataset(str(f), 10, rank=1, world=2, shuffle=False)
    assert d0.windows == 8
    x0, y0 = d0.next_batch(2)
    x1, _ = d1.next_batch(2)
    assert x0[0, 0] == 0 and x1[0, 0] == 10 and x0[1, 0] == 20 and x1[1, 0] == 30  # interleaved windows, no overlap
    assert np.array_equal(y0, x0 + 1)
    sd = d0.state_dict()
    nxt = d0.next_batch(1)[0]
    d0.load_state_dict(sd)
    assert np.array_equal(d0.next_batch(1)[0], nxt)
    # two files concatenate; wrap-around past the end
    g = tmp_path / "b.bin"
    np.arange(1000, 1000 + 2 * 10 + 1, dtype=np.uint16).tofile(g)
    d = MemmapTokenDataset(f"{f},{g}", 10, shuffle=False)
    assert d.windows == 10
    firsts = [int(d.next_batch(1)[0][0, 0]) for _ in range(11)]
    assert firsts[8] == 1000 and firsts[9] == 1010 and firsts[10] == 0


def test_memmap_dataset_shuffle_is_a_seeded_permutation_per_epoch(tmp_path):
    n = 37  # windows (prime: every multiplier is co-prime) of seq 4
    np.arange(n * 4 + 1, dtype=np.uint32).tofile(tmp_path / "t.bin")
    mk = lambda r, seed=5: MemmapTokenDataset(str(tmp_path / "t.bin"), 4, rank=r, world=2, dtype=np.uint32, seed=seed)  # noqa: E731
    d0, d1 = mk(0), mk(1)
    firsts = []
    for _ in range(n):  # 2 ranks x n draws = two epochs
        firsts += [int(d0.next_batch(1)[0][0, 0]) // 4, int(d1.next_batch(1)[0][0, 0]) // 4]
    e0, e1 = firsts[:n], firsts[n : 2 * n]
    assert sorted(e0) == list(range(n)) and sorted(e1) == list(range(n))  # every window exactly once per epoch, across ranks
    assert e0 != list(range(n)) and e0 != e1  # shuffled, and differently in the next epoch
    assert d0.epoch == 2  # both epochs are used up
    again = mk(0)
    assert int(again.next_batch(1)[0][0, 0]) // 4 == e0[0]  # same seed → same order
    assert int(mk(0, seed=6).next_batch(3)[0][2, 0]) // 4 != e0[4] or int(mk(0, seed=6).next_batch(1)[0][0, 0]) // 4 != e0[0]
    # the position is one integer
    sd = again.state_dict()
    nxt = again.next_batch(2)[0]
    again.load_state_dict(sd)
    assert np.array_equal(again.next_batch(2)[0], nxt)


def test_memmap_dataset_paths_dtypes_and_vocab_check(tmp_path):
    import json

    (tmp_path / "shards").mkdir()
    np.arange(0, 41, dtype=np.uint16).tofile(tmp_path / "shards" / "a.bin")
    np.save(tmp_path / "shards" / "b.npy", np.arange(100, 141, dtype=np.int32))
    big = tmp_path / "shards" / "c.bin"
    np.arange(70000, 70041, dtype=np.uint32).tofile(big)
    big.with_suffix(".meta.json").write_text(json.dumps({"dtype": "uint32", "vocab_size": 100000, "n_tokens": 41}))
    d = MemmapTokenDataset(str(tmp_path / "shards"), 10, vocab_size=100000, shuffle=False, dtype=np.uint16)  # sidecar overrides the dtype
    assert [f.name for f in d.files] == ["a.bin", "b.npy", "c.bin"] and d.windows == 12
    firsts = [int(d.next_batch(1)[0][0, 0]) for _ in range(12)]
    assert firsts[:4] == [0, 10, 20, 30] and firsts[4:8] == [100, 110, 120, 130] and firsts[8] == 70000
    g = MemmapTokenDataset(str(tmp_path / "shards" / "*.bin"), 10, vocab_size=100000, shuffle=False, dtype=np.uint16)
    assert [f.name for f in g.files] == ["a.bin", "c.bin"]
    with pytest.raises(ValueError, match="vocab_size"):
        MemmapTokenDataset(str(big), 10, vocab_size=32000).next_batch(1)
    with pytest.raises(ValueError, match="token_dtype"):
        MemmapTokenDataset(str(tmp_path / "shards" / "a.bin"), 10, vocab_size=100000)  # 82 bytes are not uint32 tokens
    with pytest.raises(FileNotFoundError):
        MemmapTokenDataset(str(tmp_path / "nothing*.bin"), 10)
    with pytest.raises(ValueError, match="fewer"):
        MemmapTokenDataset(str(tmp_path / "shards" / "a.bin"), 10, world=8, rank=0)


def test_prefetcher_state_is_the_consumer_position():
    ds = FakeTokenDataset(500, 8, seed=1)
    pf = PinnedPrefetcher(ds, 2, torch.device("cpu"), depth=3)
    first = pf.next()
    sd = pf.state_dict()  # two more batches are already drawn from the dataset, none of them consumed
    expect = [pf.next().input_ids.clone() for _ in range(3)]
    ds2 = FakeTokenDataset(500, 8, seed=1)
    pf2 = PinnedPrefetcher(ds2, 2, torch.device("cpu"), depth=3)
    pf2.load_state_dict(sd)
    got = [pf2.next().input_ids.clone() for _ in range(3)]
    assert all(torch.equal(a, b) for a, b in zip(expect, got)) and not torch.equal(first.input_ids, expect[0])


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [2, 4])
def test_prefetcher_no_host_dma_race_when_host_runs_ahead(depth):
    """The CPU enqueues many batches while the GPU is still busy with a long kernel (what CUDA graphs / log_interval > 1 do):
    every batch the device sees must be exactly the dataset's next batch, inputs and labels of the same draw."""
    dev = torch.device("cuda", 0)
    V, S, B, N = 30000, 512, 8, 24
    pf = PinnedPrefetcher(FakeTokenDataset(V, S, seed=5), B, dev, depth=depth)
    oracle = FakeTokenDataset(V, S, seed=5)
    sums_x = torch.zeros(N, dtype=torch.int64, device=dev)
    sums_y = torch.zeros(N, dtype=torch.int64, device=dev)
    first_x = torch.zeros(N, dtype=torch.int64, device=dev)
    a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    for i in range(N):
        if i % 3 == 0:  # keep the GPU far behind the host
            for _ in range(6):
                a = (a @ a).clamp_(-1, 1)
        b = pf.next()
        sums_x[i] = b.input_ids.sum()
        sums_y[i] = b.labels.sum()
        first_x[i] = b.input_ids[0, 0]
    torch.cuda.synchronize()
    for i in range(N):
        x, y = oracle.next_batch(B)
        assert int(sums_x[i]) == int(x.sum()) and int(sums_y[i]) == int(y.sum()) and int(first_x[i]) == int(x[0, 0]), f"batch {i} torn"


def test_tokenize_corpus_to_training_run(tmp_path):
    """text → shards (tools/tokenize_corpus.py, byte tokenizer) → MemmapTokenDataset → a real (CPU) training run on them."""
    import json
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from tools.tokenize_corpus import main as tokenize

    from prime_b200.config import load_config
    from prime_b200.train import train

    (tmp_path / "a.txt").write_text("the quick brown fox jumps over the lazy dog. " * 40)
    (tmp_path / "b.jsonl").write_text("\n".join(json.dumps({"text": f"document number {i} says hello world " * 10}) for i in range(6)))
    s = tokenize([str(tmp_path / "a.txt"), str(tmp_path / "b.jsonl"), "--tokenizer", "bytes", "--out", str(tmp_path / "ds"), "--shard-tokens", "2000"])
    assert s["documents"] == 7 and s["dtype"] == "uint16" and len(s["shards"]) == -(-s["tokens"] // 2000)
    raw = np.fromfile(tmp_path / "ds" / "shard_00000.bin", dtype=np.uint16)
    assert bytes(raw[:9].astype(np.uint8)).decode() == "the quick" and int((np.concatenate([np.fromfile(tmp_path / "ds" / f, dtype=np.uint16) for f in s["shards"]]) == 256).sum()) == 7
    d = MemmapTokenDataset(str(tmp_path / "ds"), 32, vocab_size=257)
    assert d.windows == sum((json.loads((tmp_path / "ds" / f).with_suffix(".meta.json").read_text())["n_tokens"] - 1) // 32 for f in s["shards"])
    out = train(load_config(["--name_model", "debugmodel", "--data.fake", "false", "--data.dataset_name_or_paths", str(tmp_path / "ds"), "--data.seq_length", "32",
                             "--optim.batch_size", "4", "--train.micro_bs", "2", "--optim.warmup_steps", "2", "--optim.total_steps", "12",
                             "--optim.optim.lr", "3e-3", "--diloco.inner_steps", "4", "--monitor.jsonl_path", str(tmp_path / "log.jsonl")]))  # fmt: skip
    rows = [json.loads(x) for x in (tmp_path / "log.jsonl").read_text().splitlines()]
    assert out["step"] == 12 and rows[-1]["loss"] < rows[0]["loss"] - 0.5  # bytes of English text are learnable


def test_eval_perplexity_and_generation_from_a_checkpoint(tmp_path):
    """Train on byte tokens, then evaluate the published checkpoint offline: a trained model beats the random one, greedy decoding
    is deterministic, sampling respects the seed."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from tools.tokenize_corpus import main as tokenize

    from prime_b200 import eval as ev
    from prime_b200.config import load_config
    from prime_b200.models.llama import build_model
    from prime_b200.train import train

    (tmp_path / "a.txt").write_text("abcdefgh " * 600)
    tokenize([str(tmp_path / "a.txt"), "--tokenizer", "bytes", "--out", str(tmp_path / "ds")])
    train(load_config(["--name_model", "debugmodel", "--data.fake", "false", "--data.dataset_name_or_paths", str(tmp_path / "ds"), "--data.seq_length", "32",
                       "--optim.batch_size", "8", "--train.micro_bs", "4", "--optim.warmup_steps", "2", "--optim.total_steps", "40", "--optim.optim.lr", "3e-3",
                       "--diloco.inner_steps", "10", "--ckpt.path", str(tmp_path / "ck"), "--ckpt.interval", "40"]))  # fmt: skip
    step = str(tmp_path / "ck" / "step_000040")
    res = ev.main(["ppl", "--ckpt", step, "--model", "debugmodel", "--data", str(tmp_path / "ds"), "--seq", "32", "--batches", "4", "--batch-size", "4", "--device", "cpu"])
    rnd = ev.perplexity(build_model("debugmodel", dtype=torch.float32, seed=1), MemmapTokenDataset(str(tmp_path / "ds"), 32, vocab_size=2048, shuffle=False),
                        batches=4, batch_size=4, device="cpu")  # fmt: skip
    assert res["tokens"] == 4 * 4 * 32 and res["loss"] < rnd["loss"] - 1.0 and res["perplexity"] < 50
    prompt = list(b"abcdefgh abc")
    g1 = ev.main(["generate", "--ckpt", step, "--model", "debugmodel", "--prompt-ids", ",".join(map(str, prompt)), "--max-new", "6", "--device", "cpu"])["ids"]
    g2 = ev.main(["generate", "--ckpt", step, "--model", "debugmodel", "--prompt-ids", ",".join(map(str, prompt)), "--max-new", "6", "--device", "cpu"])["ids"]
    assert g1 == g2 and len(g1) == len(prompt) + 6 and g1[: len(prompt)] == prompt
    assert bytes(g1[len(prompt) :]) == b"defgh "  # it learned the cycle
    s1 = ev.main(["generate", "--ckpt", step, "--model", "debugmodel", "--prompt-ids", "97,98", "--max-new", "5", "--temperature", "1.5", "--top-k", "5", "--seed", "3", "--device", "cpu"])["ids"]
    s2 = ev.main(["generate", "--ckpt", step, "--model", "debugmodel", "--prompt-ids", "97,98", "--max-new", "5", "--temperature", "1.5", "--top-k", "5", "--seed", "3", "--device", "cpu"])["ids"]
    assert s1 == s2


@pytest.mark.slow
def test_finetune_example_runs_end_to_end_on_cpu(tmp_path):
    """examples/finetune_from_hf.sh without a checkpoint: tiny donor → tokenise → train.init_weights run with validation → offline
    perplexity → HF directory that transformers loads."""
    import os
    import subprocess
    from pathlib import Path

    transformers = pytest.importorskip("transformers")
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run(["bash", str(root / "examples" / "finetune_from_hf.sh")], env={**os.environ, "OUT": str(tmp_path), "PYTHONPATH": str(root)},
                       capture_output=True, text=True, timeout=600)  # fmt: skip
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "validation loss" in r.stderr + r.stdout and '"perplexity"' in r.stdout
    m = transformers.AutoModelForCausalLM.from_pretrained(tmp_path / "hf_out", torch_dtype=torch.float32)
    assert m.config.hidden_size == 256 and m.config.num_hidden_layers == 2
