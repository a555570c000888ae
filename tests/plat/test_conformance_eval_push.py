"""Conformance of ``prime eval push``'s local half with the reference: which paths are accepted (and how a file inside the run
directory is forgiven), what the loader extracts from metadata.json / results.jsonl, which lines it skips and says so
(scenarios: packages/prime/tests/test_eval_push.py:19-360; harness is ours)."""

import json

import pytest
from typer.testing import CliRunner

from prime_b200.platform.commands.evals import has_eval_files, load_eval_directory, push_single_eval, validate_eval_path
from prime_b200.platform.main import app

runner = CliRunner()


def write_run(d, metadata=None, results="", *, meta=True, res=True):
    if meta:
        (d / "metadata.json").write_text(json.dumps(metadata or {"env": "gsm8k", "model": "gpt-4"}))
    if res:
        (d / "results.jsonl").write_text(results)
    return d


def test_a_run_directory_needs_both_files(tmp_path):
    assert not has_eval_files(tmp_path)
    write_run(tmp_path, res=False)
    assert not has_eval_files(tmp_path)
    write_run(tmp_path)
    assert has_eval_files(tmp_path) and validate_eval_path(str(tmp_path)) == tmp_path
    # pointing at either file inside a complete run directory means the directory
    assert validate_eval_path(str(tmp_path / "metadata.json")) == tmp_path and validate_eval_path(str(tmp_path / "results.jsonl")) == tmp_path


@pytest.mark.parametrize("present, wanted", [("metadata.json", "must contain both metadata.json and results.jsonl"),
                                             ("results.jsonl", "must contain both metadata.json and results.jsonl")])  # fmt: skip
def test_a_lone_file_is_refused(tmp_path, present, wanted):
    (tmp_path / present).write_text("{}")
    with pytest.raises(ValueError, match=wanted):
        validate_eval_path(str(tmp_path / present))


def test_other_bad_paths_say_what_is_wrong(tmp_path):
    (tmp_path / "random.txt").write_text("x")
    with pytest.raises(ValueError) as e:
        validate_eval_path(str(tmp_path / "random.txt"))
    assert "Expected a directory path" in str(e.value) and "random.txt" in str(e.value)
    with pytest.raises(ValueError, match="missing both metadata.json and results.jsonl"):
        validate_eval_path(str(tmp_path))
    write_run(tmp_path, res=False)
    with pytest.raises(ValueError, match="missing results.jsonl"):
        validate_eval_path(str(tmp_path))
    (tmp_path / "metadata.json").unlink()
    write_run(tmp_path, meta=False)
    with pytest.raises(ValueError, match="missing metadata.json"):
        validate_eval_path(str(tmp_path))
    with pytest.raises(FileNotFoundError, match="Path not found"):
        validate_eval_path("/nonexistent/path/to/eval")


def test_loader_extracts_names_metrics_and_samples(tmp_path):
    md = {"env": "gsm8k", "model": "gpt-4", "num_examples": 100, "avg_reward": 0.85, "avg_accuracy": 0.9}
    write_run(tmp_path, md, "\n".join(json.dumps(r) for r in ({"id": 0, "reward": 1.0}, {"id": 1, "reward": 0.7})))
    data = load_eval_directory(tmp_path)
    assert (data["eval_name"], data["model_name"], data["env"]) == ("gsm8k-gpt-4", "gpt-4", "gsm8k")
    assert data["metrics"] == {"reward": 0.85, "accuracy": 0.9} and data["metadata"]["num_examples"] == 100
    assert [r["example_id"] for r in data["results"]] == [0, 1] and data["results"][0]["id"] == 0  # `id` kept, `example_id` added


def test_loader_field_rules(tmp_path):
    write_run(tmp_path, {"env_id": "math-problems", "model": "claude-3"})
    data = load_eval_directory(tmp_path)
    assert data["eval_name"] == "math-problems-claude-3" and data["env"] == "math-problems"
    write_run(tmp_path, {"model": "gpt-4"})
    with pytest.raises(ValueError, match="env_id"):
        load_eval_directory(tmp_path)
    write_run(tmp_path, {"env": "gsm8k"})
    with pytest.raises(ValueError, match="model"):
        load_eval_directory(tmp_path)
    (tmp_path / "metadata.json").write_text("not valid json {")
    with pytest.raises(json.JSONDecodeError):
        load_eval_directory(tmp_path)


def test_loader_skips_bad_lines_and_says_how_many(tmp_path, capsys):
    write_run(tmp_path, {"env": "test", "model": "test-model"}, '{"id": 0, "reward": 1.0}\nnot valid json\n{"id": 1, "reward": 0.5}\n')
    assert len(load_eval_directory(tmp_path)["results"]) == 2
    out = capsys.readouterr().out
    assert "Warning" in out and "Skipped" in out
    write_run(tmp_path, {"env": "test", "model": "test-model"}, '{"id": 0, "reward": 1.0}\n123\nnull\ntrue\n"string"\n{"id": 1, "reward": 0.5}')
    data = load_eval_directory(tmp_path)
    assert [r["example_id"] for r in data["results"]] == [0, 1]
    out = capsys.readouterr().out
    assert "Warning" in out and "Skipped 4" in out and "expected dict" in out


def test_public_cannot_be_combined_with_an_existing_evaluation(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    write_run(tmp_path)
    r = runner.invoke(app, ["eval", "push", ".", "--eval-id", "eval-123", "--public"])
    assert r.exit_code == 1 and "cannot be used with --eval-id" in r.output


class RecordingEvals:
    def __init__(self):
        self.created = None

    def create_evaluation(self, **kw):
        self.created = kw
        return {"evaluation_id": "eval-new"}

    def push_samples(self, *a, **k):
        pass

    def finalize_evaluation(self, *a, **k):
        pass


@pytest.mark.parametrize("public", [False, True])
def test_new_evaluations_are_private_unless_asked(tmp_path, public):
    write_run(tmp_path)
    evals = RecordingEvals()
    assert push_single_eval(str(tmp_path), "gsm8k", None, None, is_public=public, evals=evals) == "eval-new"
    assert evals.created["is_public"] is public and evals.created["environments"] == [{"name": "gsm8k"}]


def test_encoded_batches_cut_where_build_batches_cuts():
    """The upload path serialises every sample once; its batches must be exactly the reference's (same size accounting)."""
    import json

    from prime_b200.platform.evals import build_batches, encode_batches

    samples = [{"example_id": i, "reward": i / 7, "completion": [{"role": "assistant", "content": "é✓" * (50 + 13 * (i % 29))}]} for i in range(400)]
    for limit in (2_000, 10_000, 2 * 1024 * 1024):
        plain, skipped_a = build_batches(samples, limit)
        encoded, skipped_b = encode_batches(samples, limit)
        assert skipped_a == skipped_b and [len(b) for b in plain] == [n for _, n in encoded]
        for batch, (body, n) in zip(plain, encoded):
            assert json.loads(body) == {"samples": batch} and len(body) <= limit and body.isascii()
    with pytest.warns(UserWarning, match="exceeds maximum payload size"):
        enc, skipped = encode_batches([{"x": "y" * 5000}, {"x": "ok"}], 1000)
    assert skipped == 1 and [n for _, n in enc] == [1]
