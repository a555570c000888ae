import pytest

from prime_b200.config import Config, deep_merge, load_config, parse_cli


def test_defaults():
    cfg = Config()
    assert cfg.name_model == "150M" and cfg.diloco is None
    assert cfg.optim.optim.lr == pytest.approx(4e-4)


def test_cli_over_toml_over_defaults(tmp_path):
    f = tmp_path / "c.toml"
    f.write_text('name_model = "1B"\n[optim]\nbatch_size = 64\n[optim.optim]\nlr = 1e-3\n[diloco]\ninner_steps = 100\n')
    cfg = load_config([f"@{f}", "--optim.optim.lr", "5e-4", "--train.micro_bs=8", "--data.fake"])
    assert cfg.name_model == "1B"
    assert cfg.optim.optim.lr == pytest.approx(5e-4)  # CLI wins
    assert cfg.optim.batch_size == 64  # TOML wins over default
    assert cfg.diloco.inner_steps == 100 and cfg.diloco.outer_lr == pytest.approx(0.7)
    assert cfg.train.micro_bs == 8 and cfg.data.fake is True


def test_unknown_key_rejected(tmp_path):
    f = tmp_path / "c.toml"
    f.write_text("[optim]\nbogus = 1\n")
    with pytest.raises(SystemExit) as e:
        load_config([f"@{f}"])
    assert "optim.bogus" in str(e.value)


def test_missing_file():
    with pytest.raises(SystemExit):
        load_config(["@/nonexistent/x.toml"])


def test_parse_cli_forms():
    files, ov = parse_cli(["--a.b", "3", "--c=x", "--flag", "--no-d.e"])
    assert ov == {"a": {"b": 3}, "c": "x", "flag": True, "d": {"e": False}}
    assert files == []


def test_deep_merge():
    assert deep_merge({"a": {"b": 1, "c": 2}}, {"a": {"c": 3}}) == {"a": {"b": 1, "c": 3}}


def test_ckpt_interval_must_align():
    with pytest.raises(Exception):
        Config.model_validate({"ckpt": {"interval": 7}, "diloco": {"inner_steps": 5}})
