import torch

from prime_b200.ops import reference as R


def test_int8_roundtrip_error_bound():
    x = torch.randn(5000) * 0.01
    q, s = R.quantize_int8_blockwise(x, 1024)
    y = R.dequantize_int8_blockwise(q, s, 1024)
    assert q.dtype == torch.int8 and s.numel() == 5
    # error at most half a quantisation step of the block
    step = s.repeat_interleave(1024)[:5000]
    assert torch.all((x - y).abs() <= step * 0.5 + 1e-12)


def test_int8_zero_block():
    x = torch.zeros(2048)
    q, s = R.quantize_int8_blockwise(x)
    assert torch.all(q == 0) and torch.all(s == 0)
    assert torch.all(R.dequantize_int8_blockwise(q, s) == 0)


def test_adamw_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(1000)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(1000)
        ref.grad = g.clone()
        opt.step()
        R.adamw_step(p, g, m, v, lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.1, step=step)
    torch.testing.assert_close(p, ref.detach(), rtol=1e-5, atol=1e-6)


def test_nesterov_matches_torch_sgd():
    torch.manual_seed(0)
    t = torch.randn(100)
    ref = t.clone().requires_grad_(True)
    opt = torch.optim.SGD([ref], lr=0.7, momentum=0.9, nesterov=True)
    mom = torch.zeros_like(t)
    for _ in range(4):
        g = torch.randn(100)
        ref.grad = g.clone()
        opt.step()
        R.nesterov_outer_step(t, g, mom, lr=0.7, momentum=0.9, nesterov=True)
    torch.testing.assert_close(t, ref.detach(), rtol=1e-5, atol=1e-6)


def test_rope_is_rotation():
    cos, sin = R.rope_tables(16, 8)
    x = torch.randn(2, 16, 3, 8)
    y = R.rope(x, cos, sin)
    torch.testing.assert_close(x.norm(dim=-1), y.norm(dim=-1), rtol=1e-4, atol=1e-5)
    # inverse rotation restores the input
    torch.testing.assert_close(R.rope(y, cos, -sin), x, rtol=1e-4, atol=1e-5)


def test_attention_causality():
    torch.manual_seed(0)
    q, k, v = (torch.randn(1, 8, 2, 4) for _ in range(3))
    o1 = R.attention(q, k, v, causal=True)
    k2, v2 = k.clone(), v.clone()
    k2[:, 5:], v2[:, 5:] = 0, 0  # future tokens must not influence positions < 5
    o2 = R.attention(q, k2, v2, causal=True)
    torch.testing.assert_close(o1[:, :5], o2[:, :5])
