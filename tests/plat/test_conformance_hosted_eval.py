"""Behavioural conformance of hosted evaluations (``prime eval run --hosted``, ``stop``, status printing) with the reference
CLI, observed at the HTTP boundary: what is POSTed to /hosted-evaluations and what the user is told
(scenarios: packages/prime/tests/test_hosted_eval.py:99-320, 377-460, 671-733, 979-1160; harness and fake hub are ours)."""

import pytest
from typer.testing import CliRunner

from prime_b200.platform.commands import evals as evals_mod
from prime_b200.platform.core import client as core_client
from prime_b200.platform.main import app

runner = CliRunner()


class Hub:
    def __init__(self, answer=None):
        self.posts, self.patches, self.answer = [], [], answer or {"evaluation_id": "eval-123"}

    def get(self, endpoint, params=None, **kw):
        if endpoint.startswith("/environmentshub/"):
            owner, name = endpoint.split("/")[2:4]
            return {"data": {"id": f"env-{name}", "name": name, "owner": {"name": owner}}}
        return {"data": {}}

    def post(self, endpoint, json=None, **kw):
        self.posts.append((endpoint, json))
        return self.answer

    def patch(self, endpoint, json=None, **kw):
        self.patches.append(endpoint)
        return {"message": "Evaluation cancelled", "evaluation_id": "eval-123"}


@pytest.fixture
def hub(monkeypatch):
    def install(answer=None, team=None):
        h = Hub(answer)
        monkeypatch.setenv("PRIME_API_KEY", "test-key")
        if team:
            monkeypatch.setenv("PRIME_TEAM_ID", team)
        for verb in ("get", "post", "patch"):
            monkeypatch.setattr(core_client.APIClient, verb, lambda self, *a, _v=verb, **kw: getattr(h, _v)(*a, **kw))
        return h

    return install


def test_run_hosted_defaults_and_next_step_hint(hub):
    h = hub()
    r = runner.invoke(app, ["eval", "run", "primeintellect/gsm8k", "--hosted", "-m", "openai/gpt-4.1-mini"])
    assert r.exit_code == 0, r.output
    endpoint, body = h.posts[-1]
    assert endpoint == "/hosted-evaluations" and body["environment_ids"] == ["env-gsm8k"] and body["inference_model"] == "openai/gpt-4.1-mini"
    assert body["eval_config"]["num_examples"] == 5 and body["eval_config"]["rollouts_per_example"] == 3 and "team_id" not in body
    assert "Hosted evaluation started" in r.output and "prime eval logs eval-123 -f" in r.output


def test_team_scope_and_optional_fields_reach_the_payload(hub):
    h = hub(team="cmf0ohr9s0026ilerf3w68s6n")
    r = runner.invoke(app, ["eval", "run", "primeintellect/gsm8k", "--hosted", "-m", "m", "--api-base-url", "https://llm.example/v1", "--api-key-var", "MY_KEY",
                            "--sampling-args", '{"temperature": 0.2, "extra_body": {"provider": {"order": ["azure"]}}}'])  # fmt: skip
    assert r.exit_code == 0, r.output
    body = h.posts[-1][1]
    assert body["team_id"] == "cmf0ohr9s0026ilerf3w68s6n"
    cfg = body["eval_config"]
    assert cfg["api_base_url"] == "https://llm.example/v1" and cfg["api_key_var"] == "MY_KEY"
    assert cfg["sampling_args"] == {"temperature": 0.2, "extra_body": {"provider": {"order": ["azure"]}}}


def test_plural_ids_answer_is_accepted(hub):
    hub(answer={"evaluation_ids": ["eval-123", "eval-456"]})
    r = runner.invoke(app, ["eval", "run", "primeintellect/gsm8k", "--hosted", "-m", "m"])
    assert r.exit_code == 0 and "eval-123" in r.output, r.output


def test_single_eval_toml(hub, tmp_path):
    h = hub()
    cfg = tmp_path / "eval.toml"
    cfg.write_text('''
model = "openai/gpt-4.1-mini"
num_examples = 7
rollouts_per_example = 2
timeout_minutes = 180
allow_sandbox_access = true
allow_instances_access = true
eval_name = "math500 smoke test"
sampling_args = { extra_body = { provider = { order = ["azure"], allow_fallbacks = false, require_parameters = true } } }

[[eval]]
env_id = "primeintellect/gsm8k"
env_args = { split = "test" }
''')
    r = runner.invoke(app, ["eval", "run", str(cfg), "--hosted"])
    assert r.exit_code == 0, r.output
    body = h.posts[-1][1]
    assert body["inference_model"] == "openai/gpt-4.1-mini" and body["name"] == "math500 smoke test" and body["environment_ids"] == ["env-gsm8k"]
    assert body["eval_config"] == {
        "num_examples": 7, "rollouts_per_example": 2, "timeout_minutes": 180, "allow_sandbox_access": True, "allow_instances_access": True,
        "env_args": {"split": "test"},
        "sampling_args": {"extra_body": {"provider": {"order": ["azure"], "allow_fallbacks": False, "require_parameters": True}}},
    }  # fmt: skip


def test_cli_flags_override_the_toml(hub, tmp_path):
    h = hub()
    cfg = tmp_path / "eval.toml"
    cfg.write_text('model = "from-toml"\nnum_examples = 7\n\n[[eval]]\nenv_id = "primeintellect/gsm8k"\n')
    r = runner.invoke(app, ["eval", "run", str(cfg), "--hosted", "-m", "from-cli", "-n", "11"])
    assert r.exit_code == 0, r.output
    body = h.posts[-1][1]
    assert body["inference_model"] == "from-cli" and body["eval_config"]["num_examples"] == 11


def test_multi_eval_toml_groups_identical_settings_and_splits_different_ones(hub, tmp_path):
    h = hub()
    shared = tmp_path / "shared.toml"
    shared.write_text('model = "m"\n\n[[eval]]\nenv_id = "primeintellect/gsm8k"\n\n[[eval]]\nenv_id = "primeintellect/math500"\n')
    assert runner.invoke(app, ["eval", "run", str(shared), "--hosted"]).exit_code == 0
    assert [b["environment_ids"] for _, b in h.posts] == [["env-gsm8k", "env-math500"]]  # one request for both
    h.posts.clear()
    split = tmp_path / "split.toml"
    split.write_text('model = "m"\n\n[[eval]]\nenv_id = "primeintellect/gsm8k"\nnum_examples = 3\n\n[[eval]]\nenv_id = "primeintellect/math500"\nnum_examples = 9\n')
    assert runner.invoke(app, ["eval", "run", str(split), "--hosted"]).exit_code == 0
    assert [(b["environment_ids"], b["eval_config"]["num_examples"]) for _, b in h.posts] == [(["env-gsm8k"], 3), (["env-math500"], 9)]


@pytest.mark.parametrize("toml, complaints", [
    ('timeout_minutes = "180"\n[[eval]]\nenv_id = "gsm8k"\n', ["`timeout_minutes` must be an integer"]),  # wrong type
    ('sampling_args = { until = 1979-05-27T07:32:00Z }\n[[eval]]\nenv_id = "gsm8k"\n', ["`sampling_args`", "JSON-serializable"]),
    ('resume = true\n[[eval]]\nenv_id = "gsm8k"\n', ["does not support", "`resume`"]),  # a local-only field
])  # fmt: skip
def test_toml_fields_are_checked_before_anything_is_sent(hub, tmp_path, toml, complaints):
    h = hub()
    cfg = tmp_path / "bad.toml"
    cfg.write_text(toml)
    r = runner.invoke(app, ["eval", "run", str(cfg), "--hosted"])
    assert r.exit_code == 1 and all(c in r.output for c in complaints) and not h.posts, r.output


@pytest.mark.parametrize("flag", [("--follow",), ("--poll-interval", "10")])
def test_hosted_only_flags_need_hosted(flag):
    r = runner.invoke(app, ["eval", "run", "gsm8k", *flag])
    assert r.exit_code == 1 and "hosted-only options require `--hosted`" in r.output


def test_stop_cancels_and_says_where_to_look(hub):
    h = hub()
    r = runner.invoke(app, ["eval", "stop", "eval-123"])
    assert r.exit_code == 0 and h.patches == ["/hosted-evaluations/eval-123/cancel"]
    assert "Evaluation cancelled" in r.output and "dashboard/evaluations/eval-123" in r.output


def test_status_printer_prefers_the_servers_viewer_url(monkeypatch, capsys):
    monkeypatch.setattr(evals_mod, "get_eval_viewer_url", lambda eval_id: f"fallback/{eval_id}")
    evals_mod.print_eval_status({"status": "RUNNING", "evaluation_id": "eval-123", "viewer_url": "http://localhost:3000/dashboard/evaluations/eval-123"})
    out = capsys.readouterr().out
    assert "http://localhost:3000/dashboard/evaluations/eval-123" in out and "fallback/" not in out
    evals_mod.print_eval_status({"status": "CANCELLED", "evaluation_id": "eval-123", "viewer_url": None, "error_message": "Stopped"})
    out = capsys.readouterr().out
    assert "Status: CANCELLED" in out and "Error:" in out and "fallback/eval-123" in out
