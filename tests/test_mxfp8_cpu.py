"""MXFP8 (block-scaled e4m3) — the format logic on CPU: the torch reference IS the specification the CUDA kernels are tested against
(tests/test_kernels_gpu.py compares them bit for bit), so its own properties are pinned here."""

import torch

from prime_b200 import ops
from prime_b200.ops import reference as R


def test_scales_are_smallest_power_of_two_that_fits():
    torch.manual_seed(0)
    x = (torch.randn(256, 256) * torch.exp2(torch.randint(-6, 7, (256, 1)).float())).to(torch.bfloat16)
    q, sf = R.quantize_mxfp8(x)
    idx = R._mxfp8_sf_index(256, 256)
    e = sf[idx.reshape(-1)].view(256, 8).int() - 127
    amax = x.float().view(256, 8, 32).abs().amax(-1)
    scaled = amax / torch.exp2(e.float())
    assert (scaled <= 448.0).all()          # nothing saturates
    assert (scaled[amax > 0] > 112.0).all()  # and no block wastes more than two binades of the e4m3 range
    deq = R.dequantize_mxfp8(q, sf)
    rel = (deq - x.float()).abs() / x.float().abs().clamp_min(1e-30)
    big = x.float().abs() >= (amax / 16).repeat_interleave(32, dim=1)  # elements within 4 binades of the block max keep 3 mantissa bits
    assert rel[big].max() <= 2.0**-4 + 1e-6


def test_scale_layout_is_the_tensor_core_atom():
    # [K/128][ceil(R/128)+1] atoms of 512 B; inside: (r % 32) * 16 + ((r % 128) // 32) * 4 + (kb % 4)
    idx = R._mxfp8_sf_index(300, 256)
    atoms = 300 // 128 + 1 + 1
    assert int(idx[0, 0]) == 0 and int(idx[1, 0]) == 16 and int(idx[32, 0]) == 4 and int(idx[0, 1]) == 1
    assert int(idx[128, 0]) == 512 and int(idx[0, 4]) == atoms * 512
    assert idx.unique().numel() == idx.numel() and int(idx.max()) < ops.mxfp8_sf_bytes(300, 256)


def test_transposed_quantisation_blocks_run_along_rows():
    torch.manual_seed(1)
    x = torch.randn(256, 96).to(torch.bfloat16)
    q, sf = ops.quantize_mxfp8(x, transpose=True)
    assert q.shape == (96, 256)
    assert torch.allclose(R.dequantize_mxfp8(q, sf), x.float().t(), rtol=0.07, atol=1e-3)


def test_linear_mxfp8_autograd_cpu():
    torch.manual_seed(2)
    x = torch.randn(64, 256, dtype=torch.bfloat16, requires_grad=True)
    w = (torch.randn(384, 256) * 0.05).to(torch.bfloat16).requires_grad_(True)
    y = ops.linear_mxfp8(x, w)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    (xr @ wr.t()).backward(dy.float())
    rel = lambda a, b: float((a.detach().float() - b.detach()).norm() / b.detach().norm())
    assert rel(y, xr @ wr.t()) < 0.05 and rel(x.grad, xr.grad) < 0.05 and rel(w.grad, wr.grad) < 0.02


def test_shapes_the_format_cannot_express_fall_back_to_bf16():
    x = torch.randn(8, 100, dtype=torch.bfloat16)
    w = torch.randn(60, 100, dtype=torch.bfloat16)
    assert torch.equal(ops.linear_mxfp8(x, w), ops.linear(x, w))


def test_model_fp8_switch_changes_only_block_projections():
    from prime_b200.models.llama import build_model

    model = build_model("150M", "llama2", device="cpu", dtype=torch.bfloat16, seed=0, n_layers=1)
    tokens = torch.randint(0, model.args.vocab_size, (1, 128))
    targets = torch.randint(0, model.args.vocab_size, (1, 128))
    base = float(model.loss(tokens, targets))
    model.set_fp8(True)
    assert all(b.attention.fp8 and b.feed_forward.fp8 for b in model.layers)
    fp8 = float(model.loss(tokens, targets))
    assert fp8 != base and abs(fp8 - base) < 0.02 * abs(base)
