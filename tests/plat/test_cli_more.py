"""Second batch of command-flow tests: sandbox, rl, eval, deployments, inference, images, upgrade, tunnel, MCP tools
(reference tests: packages/prime/tests/test_sandbox_*.py, test_rl_*.py, test_eval_*.py, test_upgrade.py,
packages/prime-tunnel/tests, packages/prime-mcp-server/tests)."""

import json
from datetime import datetime, timezone

import pytest
from typer.testing import CliRunner

from prime_b200.platform.commands import evals as evals_mod
from prime_b200.platform.commands import rl as rl_mod
from prime_b200.platform.commands import sandbox as sb_mod
from prime_b200.platform.core import Config
from prime_b200.platform.main import app
from prime_b200.platform.sandboxes import BulkDeleteSandboxResponse, CommandResponse, Sandbox, SandboxListResponse

runner = CliRunner()
NOW = datetime(2026, 1, 2, 3, 4, 5, tzinfo=timezone.utc)


def mk_sandbox(i, user="u1", status="RUNNING", **kw):
    return Sandbox(id=f"sb{i}", name=f"box{i}", docker_image="python:3.12-slim", cpu_cores=1, memory_gb=2, disk_size_gb=10, disk_mount_path="/data",
                   gpu_count=0, status=status, timeout_minutes=60, created_at=NOW, updated_at=NOW, user_id=user, labels=["a"], **kw)  # fmt: skip


class FakeSandboxes:
    def __init__(self, boxes):
        self.boxes, self.calls = list(boxes), []

    def list(self, **kw):
        self.calls.append(("list", kw))
        page, per = kw.get("page", 1), kw.get("per_page", 50)
        chunk = self.boxes[(page - 1) * per : page * per]
        return SandboxListResponse(sandboxes=chunk, total=len(self.boxes), page=page, per_page=per, has_next=page * per < len(self.boxes))

    def get(self, sid):
        return next(b for b in self.boxes if b.id == sid)

    def create(self, req):
        self.calls.append(("create", req))
        return mk_sandbox(99)

    def delete(self, sid):
        self.calls.append(("delete", sid))
        return {}

    def bulk_delete(self, sandbox_ids=None, labels=None):
        self.calls.append(("bulk", sandbox_ids, labels))
        ids = sandbox_ids or ["sbL"]
        return BulkDeleteSandboxResponse(succeeded=[i for i in ids if i != "bad"], failed=[{"sandbox_id": "bad", "error": "nope"}] if "bad" in ids else [], message="done")

    def execute_command(self, sid, cmd, **kw):
        self.calls.append(("exec", sid, cmd, kw))
        return CommandResponse(stdout="out\n", stderr="warn\n" if "fail" in cmd else "", exit_code=7 if "fail" in cmd else 0)

    def get_logs(self, sid):
        return "log line"


@pytest.fixture
def sandboxes(monkeypatch):
    fake = FakeSandboxes([mk_sandbox(1), mk_sandbox(2, user="u2"), mk_sandbox(3)])
    monkeypatch.setattr(sb_mod, "client", lambda: fake)
    return fake


def test_sandbox_list_get_run(sandboxes):
    out = json.loads(runner.invoke(app, ["sandbox", "list", "-o", "json"]).output)
    assert out["total"] == 3 and out["sandboxes"][0]["id"] == "sb1"
    assert "box2" in runner.invoke(app, ["sandbox", "ls"]).output  # hidden alias
    assert json.loads(runner.invoke(app, ["sandbox", "get", "sb2", "-o", "json"]).output)["id"] == "sb2"
    r = runner.invoke(app, ["sandbox", "run", "sb1", "-e", "A=1", "-w", "/tmp", "--", "echo", "hi"])
    assert r.exit_code == 0 and "out" in r.output
    assert sandboxes.calls[-1] == ("exec", "sb1", "echo hi", {"working_dir": "/tmp", "env": {"A": "1"}, "timeout": None})
    r = runner.invoke(app, ["sandbox", "run", "sb1", "fail now"])
    assert r.exit_code == 7 and "stderr:" in r.output  # the command's exit code is propagated
    assert "log line" in runner.invoke(app, ["sandbox", "logs", "sb1"]).output


def test_sandbox_create_validation_and_secrets_hidden(sandboxes):
    assert "GPU type is required" in runner.invoke(app, ["sandbox", "create", "img", "--gpu-count", "1", "--vm"]).output
    assert "require VM" in runner.invoke(app, ["sandbox", "create", "img", "--gpu-count", "1", "--gpu-type", "H100_80GB"]).output
    r = runner.invoke(app, ["sandbox", "create", "python:3.12-slim", "-e", "MODE=fast", "--secret", "TOKEN=supersecretvalue", "-l", "ci", "-y"])
    assert r.exit_code == 0 and "Successfully created sandbox sb99" in r.output and "supersecretvalue" not in r.output
    req = sandboxes.calls[-1][1]
    assert req.secrets == {"TOKEN": "supersecretvalue"} and req.environment_vars == {"MODE": "fast"} and req.labels == ["ci"] and req.name
    r = runner.invoke(app, ["sandbox", "create", "img"], input="n\n")
    assert "cancelled" in r.output and sandboxes.calls[-1][0] == "create" and sandboxes.calls[-1][1] is req  # nothing new was created


def test_sandbox_delete_modes(sandboxes, isolated_home):
    assert runner.invoke(app, ["sandbox", "delete", "sb1", "--all"]).exit_code == 1  # mutually exclusive
    assert runner.invoke(app, ["sandbox", "delete", "sb1", "-y"]).exit_code == 0 and sandboxes.calls[-1] == ("delete", "sb1")
    r = runner.invoke(app, ["sandbox", "delete", "sb1,sb2", "bad", "-y"])
    assert r.exit_code == 1 and sandboxes.calls[-1] == ("bulk", ["sb1", "sb2", "bad"], None) and "✗ bad: nope" in r.output
    assert runner.invoke(app, ["sandbox", "delete", "--label", "ci", "-y"]).exit_code == 0 and sandboxes.calls[-1] == ("bulk", None, ["ci"])
    # --all only touches the caller's sandboxes unless --all-users; needs a known user id
    assert "no user_id configured" in runner.invoke(app, ["sandbox", "delete", "--all", "-y"]).output
    Config().set_user_id("u1")
    assert runner.invoke(app, ["sandbox", "delete", "--all", "-y"]).exit_code == 0 and sandboxes.calls[-1] == ("bulk", ["sb1", "sb3"], None)
    assert runner.invoke(app, ["sandbox", "delete", "--all", "--all-users", "-y"]).exit_code == 0 and sandboxes.calls[-1][1] == ["sb1", "sb2", "sb3"]


# ----------------------------------------------------------------------------------------------- rl
RUN = {"id": "r1", "userId": "u1", "status": "QUEUED", "rolloutsPerExample": 8, "seqLen": 4096, "maxSteps": 10, "batchSize": 128, "baseModel": "m",
       "runsAhead": 2, "createdAt": "2026-01-02T03:04:05Z", "updatedAt": "2026-01-02T03:04:05Z"}  # fmt: skip


def test_rl_init_validate_and_run(tmp_path, fake_api, monkeypatch):
    cfgp = tmp_path / "rl.toml"
    r = runner.invoke(app, ["rl", "init", str(cfgp)])
    assert r.exit_code == 0 and cfgp.exists()
    cfg = rl_mod.load_config(str(cfgp))
    assert cfg.env[0].id == "primeintellect/reverse-text" and cfg.model == "PrimeIntellect/Qwen3-0.6B-Reverse-Text-SFT" and cfg.sampling.max_tokens == 2048
    bad = tmp_path / "bad.toml"
    bad.write_text('model = "m"\nbogus_key = 1\n[[env]]\nid = "a/b"\n')
    r = runner.invoke(app, ["rl", "run", str(bad)])
    assert r.exit_code == 1 and "bogus_key" in r.output
    # W&B configured without a key → refused before any API call
    wb = tmp_path / "wb.toml"
    wb.write_text('model = "m"\n[[env]]\nid = "a/b"\n[wandb]\nproject = "p"\n')
    assert "WANDB_API_KEY is required" in runner.invoke(app, ["rl", "run", str(wb)]).output
    (tmp_path / "s.env").write_text("WANDB_API_KEY=abc\nHF=${HOME_TOKEN}\n")
    monkeypatch.setenv("HOME_TOKEN", "hf_x")
    api = fake_api({("POST", "/rft/runs"): {"run": RUN}, ("GET", "/environmentshub/a/b/status"): {"data": {"latest_action": {"status": "FAILED", "error": "tests red"}}}}, rl_mod)
    r = runner.invoke(app, ["rl", "run", str(wb), "--env-file", str(tmp_path / "s.env")])
    assert r.exit_code == 1 and "latest action FAILED" in r.output and not api.called("POST", "/rft/runs")
    r = runner.invoke(app, ["rl", "run", str(wb), "--env-file", str(tmp_path / "s.env"), "--skip-action-check", "-o", "json"])
    assert r.exit_code == 0, r.output
    body = api.called("POST", "/rft/runs")[0][3]
    assert {d["key"]: d["value"] for d in body["secrets"]} == {"WANDB_API_KEY": "abc", "HF": "hf_x"} and body["environments"][0]["id"] == "a/b"
    assert '"runs_ahead": 2' in r.output and "abc" not in r.output.split("{", 1)[0]  # secrets never echoed in the summary


def test_rl_log_cleaning():
    lines = rl_mod.clean_logs('\x1b[32m{"timestamp":"2026-01-02T03:04:05Z","level":"info","message":"step 3"}\x1b[0m\n  50%|█████     | 5/10 [00:01<00:01]\nplain\n\n')
    assert any("step 3" in ln for ln in lines) and "plain" in lines and not any("█" in ln for ln in lines)
    assert rl_mod.is_queued_404(Exception("HTTP 404: run is queued")) or not rl_mod.is_queued_404(Exception("HTTP 500"))


# ----------------------------------------------------------------------------------------------- eval
def test_eval_list_get_and_push_discovery(tmp_path, fake_api, monkeypatch):
    fake_api({("GET", "/evaluations/"): {"evaluations": [{"evaluation_id": "e1", "name": "n" * 50, "model_name": "m", "status": "COMPLETED", "total_samples": 4,
                                                           "created_at": "2026-01-02T03:04:05Z"}], "total": 1},
              ("GET", "/evaluations/e1"): {"evaluation_id": "e1", "status": "COMPLETED"},
              ("GET", "/evaluations/e1/samples"): {"samples": [{"example_id": 0, "reward": 1.0, "answer": "4"}], "total": 1}}, evals_mod)  # fmt: skip
    r = runner.invoke(app, ["eval", "list"])
    assert r.exit_code == 0 and "e1" in r.output and "n" * 31 not in r.output  # long names are clipped
    assert json.loads(runner.invoke(app, ["eval", "get", "e1"]).output)["status"] == "COMPLETED"
    assert "reward=1.0" in runner.invoke(app, ["eval", "samples", "e1", "-o", "pretty"]).output
    # output-directory discovery used by `prime eval push` without arguments
    run = tmp_path / "outputs" / "evals" / "gsm8k--model" / "abc123"
    run.mkdir(parents=True)
    (run / "metadata.json").write_text(json.dumps({"env": "gsm8k", "model": "model", "num_examples": 1, "rollouts_per_example": 1}))
    (run / "results.jsonl").write_text(json.dumps({"example_id": 0, "reward": 1.0}) + "\n")
    monkeypatch.chdir(tmp_path)
    assert evals_mod.has_eval_files(run) and [p.resolve() for p in evals_mod.discover_eval_outputs()] == [run]
    assert evals_mod.validate_eval_path(str(run)) == run
    with pytest.raises(Exception):
        evals_mod.validate_eval_path(str(tmp_path / "nope"))


def test_eval_hosted_config_expansion(tmp_path):
    cfg = tmp_path / "evals.toml"
    cfg.write_text('model = "m-default"\nnum_examples = 5\n[[eval]]\nenv_id = "a/b"\n[[eval]]\nenv_id = "c/d"\nmodel = "m2"\nrollouts_per_example = 3\n')
    entries = evals_mod.load_hosted_eval_configs(str(cfg))
    assert [(e["env_id"], e["model"]) for e in entries] == [("a/b", "m-default"), ("c/d", "m2")] and entries[1]["rollouts_per_example"] == 3
    assert evals_mod.parse_json_object_option('{"a": 1}', "--x") == {"a": 1}
    for bad in ("[1]", "{nope"):
        with pytest.raises(Exception):
            evals_mod.parse_json_object_option(bad, "--x")


def test_shipped_eval_results_sample_is_a_pushable_directory():
    """examples/eval_results_sample/ (this repo's counterpart of the reference's examples/verifiers_example/): the loader takes it as is,
    its metadata agrees with its rows, and the hub payload keeps every row."""
    from pathlib import Path

    from prime_b200.platform.utils import eval_push as ep

    d = Path(__file__).resolve().parents[2] / "examples" / "eval_results_sample"
    assert evals_mod.has_eval_files(d) and evals_mod.validate_eval_path(str(d)) == d
    meta = json.loads((d / "metadata.json").read_text())
    rows = ep.load_results_jsonl(d / "results.jsonl")
    assert len(rows) == meta["num_examples"] * meta["rollouts_per_example"]
    assert abs(sum(r["reward"] for r in rows) / len(rows) - meta["avg_reward"]) < 1e-3
    assert len(ep.to_hub_samples(rows)) == len(rows)
