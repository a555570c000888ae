"""Data path: synthetic and memory-mapped token datasets, rank striding, the prefetcher's consumer-side checkpoint position."""

import numpy as np
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
    d0, d1 = MemmapTokenDataset(str(f), 10, rank=0, world=2), MemmapTokenDataset(str(f), 10, rank=1, world=2)
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
    d = MemmapTokenDataset(f"{f},{g}", 10)
    assert d.windows == 10
    firsts = [int(d.next_batch(1)[0][0, 0]) for _ in range(11)]
    assert firsts[8] == 1000 and firsts[9] == 1010 and firsts[10] == 0


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


import pytest


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
