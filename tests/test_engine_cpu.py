"""Sharded engine on CPU: must reproduce a plain (unsharded) AdamW training run."""

import copy

import pytest
import torch

from prime_b200.models.llama import build_model
from prime_b200.parallel.fsdp import AdamHyper, ShardedEngine
from prime_b200.parallel.mesh import WorldInfo, build_mesh


def _mesh():
    return build_mesh(WorldInfo(), device=torch.device("cpu"))


def test_engine_matches_plain_adamw():
    torch.manual_seed(0)
    m_ref = build_model("debugmodel", dtype=torch.float32, seed=5)
    m_eng = copy.deepcopy(m_ref)
    hyper = AdamHyper(lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.1, max_norm=1.0)
    eng = ShardedEngine(m_eng, _mesh(), hyper, backend="collective")
    opt = torch.optim.AdamW(m_ref.parameters(), lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    V = m_ref.args.vocab_size
    for step in range(4):
        tok = torch.randint(0, V, (2, 16))
        tgt = torch.randint(0, V, (2, 16))
        # reference: two micro-batches accumulated, clip at 1.0
        opt.zero_grad()
        for i in range(2):
            (m_ref.loss(tok[i : i + 1], tgt[i : i + 1]) / 2).backward()
        torch.nn.utils.clip_grad_norm_(m_ref.parameters(), 1.0)
        opt.step()
        # engine
        eng.zero_grad()
        for i in range(2):
            eng.set_micro_step(i == 1)
            m_eng.loss(tok[i : i + 1], tgt[i : i + 1], grad_scale=0.5).backward()
            if i == 1:
                eng.finish_backward()
            else:
                eng.fold_micro_grads()
        eng.step()
    for (n1, p1), (n2, p2) in zip(m_ref.named_parameters(), m_eng.named_parameters()):
        torch.testing.assert_close(p1, p2, rtol=2e-4, atol=2e-5, msg=lambda s: f"{n1}: {s}")


def test_bucket_layout_alignment():
    m = build_model("debugmodel", dtype=torch.float32, seed=5)
    eng = ShardedEngine(m, _mesh(), AdamHyper(), backend="collective")
    names = [b.name for b in eng.buckets]
    assert names[0] == "embed" and names[-1] == "head" and "layer0" in names
    for b in eng.buckets:
        assert b.size % 1024 == 0 and b.start % 1024 == 0
        for _, p, off in b.params:
            assert off % 8 == 0
            assert p.data.data_ptr() == eng.param_flat[b.start + off :].data_ptr()
            assert p.main_grad.dtype == torch.float32


def test_boundary_hooks_fire_per_layer():
    m = build_model("debugmodel", dtype=torch.float32, seed=5)
    eng = ShardedEngine(m, _mesh(), AdamHyper(), backend="collective")
    fired = []
    orig = eng._on_bucket_ready
    eng._on_bucket_ready = lambda b: (fired.append(eng.buckets[b].name), orig(b))
    eng.zero_grad()
    eng.set_micro_step(True)
    tok = torch.randint(0, m.args.vocab_size, (1, 8))
    m.loss(tok, tok).backward()
    # backward order: head first, then layers from last to first
    assert fired == ["head", "layer1", "layer0"]


@pytest.mark.parametrize("ac", [True, 2])
def test_activation_checkpointing_is_exact(ac):
    """Re-running a block's forward inside backward must not change the loss, the gradients or how often a bucket is reduced."""
    import torch

    from prime_b200.config import Config
    from prime_b200.trainer import Trainer

    def run(ac_ckpt):
        cfg = Config.model_validate({"name_model": "debugmodel", "data": {"seq_length": 32}, "optim": {"batch_size": 4, "warmup_steps": 1},
                                     "train": {"micro_bs": 2, "ac_ckpt": ac_ckpt}})  # fmt: skip
        t = Trainer(cfg)
        losses = [float(t.inner_step().loss) for _ in range(3)]
        return losses, t.engine.master.clone(), float(t.engine.last_grad_norm)

    l0, m0, g0 = run(False)
    l1, m1, g1 = run(ac)
    assert l0 == l1 and g0 == g1 and torch.equal(m0, m1)


def test_load_state_dict_rejects_other_shard_layouts():
    m = build_model("debugmodel", dtype=torch.float32, seed=5)
    eng = ShardedEngine(m, _mesh(), AdamHyper(), backend="collective")
    sd = {k: (v.clone() if isinstance(v, torch.Tensor) else copy.deepcopy(v)) for k, v in eng.state_dict().items()}
    eng.load_state_dict(sd)  # round trip
    with pytest.raises(ValueError, match="reshard_after_forward"):
        eng.load_state_dict({**sd, "shard_params": True})  # written by a ZeRO-3 run: row-block shards, not bucket slices
    with pytest.raises(ValueError, match="fsdp_size"):
        eng.load_state_dict({**sd, "fsdp_size": 2})
    bad = [list(x) for x in sd["layout"]]
    bad[1][2] += 1024
    with pytest.raises(ValueError, match="layout"):
        eng.load_state_dict({**sd, "layout": [tuple(x) for x in bad]})
