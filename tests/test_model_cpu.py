import torch

from prime_b200.models.llama import (
    build_model,
    from_reference_state_dict,
    get_model_args,
    to_reference_state_dict,
)


def test_sizes():
    a = get_model_args("1B")
    assert (a.dim, a.n_layers, a.n_heads, a.ffn_hidden) == (2048, 18, 16, 5632)
    a7 = get_model_args("7B")
    assert (a7.dim, a7.n_layers, a7.ffn_hidden) == (4096, 32, 11008)
    assert get_model_args("150M").dim == 1024


def test_forward_shapes_and_backward():
    m = build_model("debugmodel", dtype=torch.float32, seed=1)
    tok = torch.randint(0, m.args.vocab_size, (2, 16))
    logits = m(tok)
    assert logits.shape == (2, 16, m.args.vocab_size)
    loss = m.loss(tok, tok)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())


def test_param_count_1b():
    a = get_model_args("1B")
    per_layer = (a.n_heads + 2 * a.kv_heads) * a.head_dim * a.dim + a.dim * a.dim + 3 * a.dim * a.ffn_hidden + 2 * a.dim
    total = per_layer * a.n_layers + 2 * a.vocab_size * a.dim + a.dim
    assert 1.0e9 < total < 1.2e9


def test_state_dict_interchange():
    m1 = build_model("debugmodel", dtype=torch.float32, seed=1)
    m2 = build_model("debugmodel", dtype=torch.float32, seed=2)
    sd = to_reference_state_dict(m1)
    assert "layers.0.attention.wq.weight" in sd and "layers.1.feed_forward.w3.weight" in sd
    from_reference_state_dict(m2, {k: v.clone() for k, v in sd.items()})
    tok = torch.randint(0, m1.args.vocab_size, (1, 8))
    torch.testing.assert_close(m1(tok), m2(tok))


def test_fused_residual_chain_matches_plain():
    """The (h, delta) fused add+norm chain must equal the textbook pre-norm residual block."""
    from prime_b200.ops import reference as R

    m = build_model("debugmodel", dtype=torch.float32, seed=3)
    tok = torch.randint(0, m.args.vocab_size, (1, 8))
    cos, sin = m.rope_tables(8, tok.device)
    h = m.tok_embeddings(tok)
    for layer in m.layers:
        x = R.rmsnorm(h, layer.attention_norm.weight, layer.eps)
        h = h + layer.attention(x, cos, sin)
        y = R.rmsnorm(h, layer.ffn_norm.weight, layer.eps)
        h = h + layer.feed_forward(y)
    ref = R.rmsnorm(h, m.norm.weight, m.args.norm_eps)
    torch.testing.assert_close(m.forward_hidden(tok), ref, rtol=1e-4, atol=1e-5)
