import pytest
import torch

from prime_b200.models.llama import (
    Transformer,
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


# ---------------------------------------------------------------------------------------------- Hugging Face interchange
def _tiny_hf(kv_heads: int = 2):
    transformers = pytest.importorskip("transformers")
    cfg = transformers.LlamaConfig(vocab_size=512, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=kv_heads, max_position_embeddings=128, rms_norm_eps=1e-5, rope_theta=10000.0,
                                   tie_word_embeddings=False, attn_implementation="eager")  # fmt: skip
    torch.manual_seed(3)
    return transformers.LlamaForCausalLM(cfg).eval()


@pytest.mark.parametrize("kv_heads", [4, 2])
def test_hf_import_matches_transformers_logits(kv_heads):
    """HF rotates (x[i], x[i + D/2]) and permutes q/k rows per head; this engine rotates interleaved pairs. After the import the
    two models must be the same function (MHA and GQA)."""
    from prime_b200.models import hf

    ref = _tiny_hf(kv_heads)
    args = hf.args_from_hf_config(ref.config.to_dict())
    assert (args.dim, args.n_layers, args.n_heads, args.kv_heads, args.ffn_hidden, args.vocab_size) == (64, 2, 4, kv_heads, 176, 512)
    m = Transformer(args).float()
    hf.load_hf_state_dict(m, ref.state_dict())
    tok = torch.randint(0, 512, (2, 48))
    with torch.no_grad():
        want = ref(tok).logits
        got = m(tok)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    # and back: the exported state dict is exactly what transformers had
    back = hf.to_hf_state_dict(m)
    for k, v in ref.state_dict().items():
        torch.testing.assert_close(back[k], v, rtol=0, atol=0, msg=k)


def test_hf_export_directory_loads_in_transformers(tmp_path):
    transformers = pytest.importorskip("transformers")
    pytest.importorskip("safetensors")
    from prime_b200.models import hf

    m = build_model("debugmodel", dtype=torch.float32, seed=11, n_kv_heads=4)
    names = hf.save_hf_dir(m, tmp_path / "one")
    assert names == ["model.safetensors"]
    sharded = hf.save_hf_dir(m, tmp_path / "many", max_shard_bytes=2 << 20)  # force an index + several shards
    assert len(sharded) > 1 and (tmp_path / "many" / "model.safetensors.index.json").exists()
    tok = torch.randint(0, m.args.vocab_size, (1, 40))
    with torch.no_grad():
        want = m(tok)
    for d in ("one", "many"):
        loaded = transformers.AutoModelForCausalLM.from_pretrained(tmp_path / d, torch_dtype=torch.float32, attn_implementation="eager").eval()
        with torch.no_grad():
            torch.testing.assert_close(loaded(tok).logits, want, rtol=1e-4, atol=1e-4)
        again = hf.load_hf_dir(tmp_path / d, dtype=torch.float32)  # and our own reader takes both layouts
        for (k, a), (_, b) in zip(m.named_parameters(), again.named_parameters()):
            torch.testing.assert_close(a, b, rtol=0, atol=0, msg=k)


def test_hf_config_refuses_what_it_cannot_represent():
    from prime_b200.models import hf

    base = hf.hf_config_from_args(build_model("debugmodel", dtype=torch.float32).args)
    assert hf.args_from_hf_config(base).ffn_hidden == base["intermediate_size"]
    with pytest.raises(ValueError, match="rope_scaling"):
        hf.args_from_hf_config({**base, "rope_scaling": {"rope_type": "llama3", "factor": 8.0}})
    with pytest.raises(ValueError, match="Llama"):
        hf.args_from_hf_config({**base, "model_type": "mistral", "architectures": ["MistralForCausalLM"]})
    with pytest.raises(ValueError, match="bias"):
        hf.args_from_hf_config({**base, "attention_bias": True})


def test_b0_comparison_model_is_the_same_function_as_ours():
    """The stock-PyTorch arm of the benchmark (baseline/torch_b0.Llama: nn.Linear, SDPA, complex-multiply RoPE, nn.RMSNorm) must be
    the SAME network as the engine's, or the reported ratio compares different jobs. Checked here on the CPU in fp32 (the GPU suite
    repeats it at the 1B shape against the native kernels): identical logits, loss and parameter gradients for the same weights."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from baseline.torch_b0 import Llama as B0Llama

    ours = build_model("150M", dtype=torch.float32, seed=3, n_layers=2, vocab_size=4096, max_seq_len=64)
    a = ours.args
    ref = B0Llama(a.dim, a.n_layers, a.n_heads, a.vocab_size, max_seq=64).float()
    with torch.no_grad():
        ref.tok_embeddings.weight.copy_(ours.tok_embeddings.weight)
        ref.norm.weight.copy_(ours.norm.weight)
        ref.output.weight.copy_(ours.output)
        for x, y in zip(ours.layers, ref.layers):
            y.wqkv.weight.copy_(x.attention.wqkv)
            y.wo.weight.copy_(x.attention.wo)
            y.w13.weight.copy_(x.feed_forward.w13)
            y.w2.weight.copy_(x.feed_forward.w2)
            y.attention_norm.weight.copy_(x.attention_norm.weight)
            y.ffn_norm.weight.copy_(x.ffn_norm.weight)
    assert sum(p.numel() for p in ref.parameters()) == ours.num_params()
    tok = torch.randint(0, a.vocab_size, (2, 48))
    tgt = torch.randint(0, a.vocab_size, (2, 48))
    lo, lr = ours(tok), ref(tok)
    torch.testing.assert_close(lo, lr, rtol=2e-4, atol=2e-4)
    torch.nn.functional.cross_entropy(lo.reshape(-1, a.vocab_size), tgt.reshape(-1)).backward()
    torch.nn.functional.cross_entropy(lr.reshape(-1, a.vocab_size), tgt.reshape(-1)).backward()
    for x, y in zip(ours.layers, ref.layers):
        torch.testing.assert_close(x.attention.wqkv.grad, y.wqkv.weight.grad, rtol=2e-3, atol=2e-5)
        torch.testing.assert_close(x.feed_forward.w2.grad, y.w2.weight.grad, rtol=2e-3, atol=2e-5)
    torch.testing.assert_close(ours.tok_embeddings.weight.grad, ref.tok_embeddings.weight.grad, rtol=2e-3, atol=2e-5)
    assert abs(ours.flops_per_token(48) - ref.flops_per_token(48)) < 1e-6 * ref.flops_per_token(48)  # the MFU denominators agree too
