"""verifiers plugin / bridge, eval push, environment metadata, time helpers, hosted-eval log handling, Connect-RPC helpers
(reference tests: packages/prime/tests/test_verifiers_bridge.py, test_eval_push.py, test_env_metadata.py, test_time_utils.py,
test_hosted_eval.py; packages/prime-sandboxes/tests/test_rpc_command_session.py)."""

import json
import os
import sys
import types
from datetime import datetime, timedelta, timezone
from pathlib import Path

import pytest

from prime_b200.platform import verifiers_bridge as vb
from prime_b200.platform import verifiers_plugin as vp
from prime_b200.platform.utils import env_metadata as em
from prime_b200.platform.utils import eval_push as ep
from prime_b200.platform.utils import hosted_eval as he
from prime_b200.platform.utils import time_utils as tu


# ------------------------------------------------------------------------------------------------ verifiers plugin
def test_plugin_falls_back_to_builtin_mapping(monkeypatch):
    class Sink:
        def __init__(self):
            self.lines = []

        def print(self, *a, **k):
            self.lines.append(" ".join(map(str, a)))

    sink = Sink()
    monkeypatch.setitem(sys.modules, "verifiers.cli.plugins.prime", None)  # import fails
    p = vp.load_verifiers_prime_plugin(sink)
    assert p == vp.PrimeVerifiersPlugin() and "Falling back" in sink.lines[0]
    # a plugin with a different API version and one overridden module is honoured, with a warning
    mod = types.ModuleType("verifiers.cli.plugins.prime")
    mod.get_plugin = lambda: types.SimpleNamespace(api_version=99, eval_module="custom.eval")
    for name in ("verifiers", "verifiers.cli", "verifiers.cli.plugins"):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    monkeypatch.setitem(sys.modules, "verifiers.cli.plugins.prime", mod)
    sink2 = Sink()
    p = vp.load_verifiers_prime_plugin(sink2)
    assert p.eval_module == "custom.eval" and p.gepa_module == vp.PrimeVerifiersPlugin().gepa_module and p.api_version == 99
    assert "version mismatch" in sink2.lines[0]


def test_workspace_python_resolution_order(tmp_path, monkeypatch):
    ws = tmp_path / "proj" / "sub"
    ws.mkdir(parents=True)
    (tmp_path / "proj" / "pyproject.toml").write_text("[project]\nname='x'\n")
    venv_py = vp.venv_python(tmp_path / "proj" / ".venv")
    venv_py.parent.mkdir(parents=True)
    venv_py.write_text("")
    monkeypatch.delenv("UV_PROJECT_ENVIRONMENT", raising=False)
    monkeypatch.setenv("VIRTUAL_ENV", str(tmp_path / "active"))
    cands = list(vp.candidate_interpreters(ws))
    assert cands[0] == vp.venv_python(tmp_path / "active") and venv_py in cands
    vp.can_import.cache_clear()
    monkeypatch.setattr(vp, "can_import", lambda python, module, cwd: python == str(venv_py))
    assert vp.resolve_workspace_python(ws) == str(venv_py)
    monkeypatch.setattr(vp, "can_import", lambda python, module, cwd: False)
    assert vp.resolve_workspace_python(ws) == sys.executable
    assert vp.PrimeVerifiersPlugin().build_module_command("m.x", ["--a"])[1:] == ["-m", "m.x", "--a"]


# ------------------------------------------------------------------------------------------------ verifiers bridge
def test_passthrough_argv_helpers():
    assert vb.is_help_request("gsm8k", ["-n", "5", "--help"]) and vb.is_help_request("-h", []) and not vb.is_help_request("gsm8k", ["-n", "5"])
    args = ["-m", "gpt", "--num-examples=7", "-r3", "--env-dir-path", "envs"]
    assert vb.parse_value_option(args, "--model", "-m") == "gpt"
    assert vb.parse_value_option(args, "--num-examples", "-n") == "7"
    assert vb.parse_value_option(args, "--rollouts", "-r") == "3"
    assert vb.parse_value_option(args, "--missing", "-x") is None and vb.parse_value_option(["--model"], "--model", "-m") is None
    assert vb.has_flag(args, "--num-examples", "-n") and not vb.has_flag(args, "--verbose", "-v")
    assert vb.is_config_target("eval.toml") and not vb.is_config_target("gsm8k")
    assert vb.split_version("owner/env@1.2.0") == ("owner/env", "1.2.0") and vb.split_version("env") == ("env", None)
    assert vb.is_slug_reference("owner/env") and not vb.is_slug_reference("./local/env") and not vb.is_slug_reference("cfg/eval.toml")
    assert vb.split_owner_and_name("a/b") == ("a", "b") and vb.split_owner_and_name("a/") is None
    assert vb.environment_url_from_slug("a/b").endswith("/dashboard/environments/a/b") and vb.environment_url_from_slug("nope") is None
    jid = vb.build_job_id("my-env", "org/model-x")
    assert jid.startswith("my_env_org_model_x_") and len(jid.rsplit("_", 1)[1]) == 8


def make_env(root: Path, name="my_env") -> Path:
    d = root / name
    (d / "data").mkdir(parents=True)
    (d / "__pycache__").mkdir()
    (d / "pyproject.toml").write_text("[project]\nname='my-env'\nversion='0.1.0'\n")
    (d / "my_env.py").write_text("def load_environment():\n    return 1\n")
    (d / "README.md").write_text("# env\n")
    (d / "data" / "a.jsonl").write_text("{}\n")
    (d / "__pycache__" / "x.pyc").write_bytes(b"\0")
    (d / "data" / ".hidden").write_text("x")
    return d


def test_content_hash_ignores_caches_and_tracks_content(tmp_path):
    d = make_env(tmp_path)
    h1 = vb.compute_local_content_hash(d)
    assert vb.is_valid_hash(h1) and not vb.is_valid_hash("xyz") and not vb.is_valid_hash(None)
    (d / "__pycache__" / "y.pyc").write_bytes(b"1")
    (d / "data" / ".other").write_text("ignored")
    assert vb.compute_local_content_hash(d) == h1  # caches and dotfiles do not count
    (d / "data" / "a.jsonl").write_text('{"x": 1}\n')
    assert vb.compute_local_content_hash(d) != h1
    assert vb.compute_local_content_hash(tmp_path / "missing") is None
    kinds = [k for k, _ in vb.hashed_items(d)]
    assert "dir" in kinds and kinds.count("file") == 4


class HubAPI:
    """whoami / teams / environment lookups keyed by owner."""

    def __init__(self, envs, me="ada", teams=()):
        self.envs, self.me, self.teams, self.calls = envs, me, list(teams), []

    def get(self, endpoint, params=None, **kw):
        from prime_b200.platform.core import APIError

        self.calls.append(endpoint)
        if endpoint == "/user/whoami":
            return {"data": {"slug": self.me}}
        if endpoint == "/user/teams":
            return {"data": self.teams}
        if endpoint.startswith("/environmentshub/"):
            _, _, owner, name, _ver = endpoint.split("/", 4)
            if (owner, name) in self.envs:
                return {"data": self.envs[(owner, name)]}
            raise APIError("HTTP 404", 404)
        raise APIError("HTTP 404", 404)


def test_resolve_slug_local_and_remote_references(tmp_path, monkeypatch):
    from prime_b200.platform.core import Config

    r = vb.resolve_environment_reference("owner/env@0.2.0", str(tmp_path))
    assert (r.install_mode, r.install_slug, r.upstream_slug, r.env_name) == ("remote", "owner/env@0.2.0", "owner/env", "env")
    # local directory, never pushed → local only, recommend push
    d = make_env(tmp_path)
    api = HubAPI({})
    r = vb.resolve_environment_reference("my-env", str(tmp_path), client=api, config=Config(writable=False))
    assert r.install_mode == "local" and r.recommend_push and r.push_reason == "local_only" and r.local_env_path == d and r.upstream_slug is None
    # pushed and in sync (content hash matches) → tracked slug, no push recommendation
    em.write_environment_metadata(d, {"owner": "ada", "name": "my-env", "version": "0.1.0"})
    h = vb.compute_local_content_hash(d)
    api = HubAPI({("ada", "my-env"): {"latest_version": {"semantic_version": "0.1.0", "content_hash": h}}})
    r = vb.resolve_environment_reference("my-env", str(tmp_path), client=api, config=Config(writable=False))
    assert r.upstream_slug == "ada/my-env" and not r.recommend_push and r.env_display_id == "ada/my-env"
    # local edits → ahead of the hub
    (d / "my_env.py").write_text("def load_environment():\n    return 2\n")
    r = vb.resolve_environment_reference("my-env", str(tmp_path), client=api, config=Config(writable=False))
    assert r.recommend_push and r.push_reason == "ahead" and "ahead of ada/my-env" in r.env_display_id
    # not local: personal owner wins, then team, then the official namespace
    other = tmp_path / "elsewhere"
    other.mkdir()
    api = HubAPI({("ada", "gsm8k"): {"id": "e1"}})
    r = vb.resolve_environment_reference("gsm8k", str(other), client=api, config=Config(writable=False))
    assert r.install_mode == "remote" and r.install_slug == "ada/gsm8k"
    api = HubAPI({(vb.PRIME_SLUG, "gsm8k"): {"id": "e2"}})
    r = vb.resolve_environment_reference("gsm8k@1.0", str(other), client=api, config=Config(writable=False))
    assert r.install_slug == f"{vb.PRIME_SLUG}/gsm8k@1.0"
    assert vb.resolve_environment_reference("nothing", str(other), client=HubAPI({}), config=Config(writable=False)).install_mode == "none"
    assert vb.remote_version_and_hash({"semantic_version": "1", "sha256": "ab"}) == ("1", "ab") and vb.remote_version_and_hash(None) == (None, None)


# ------------------------------------------------------------------------------------------------ env metadata
def test_metadata_lookup_order_and_migration(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    legacy = tmp_path / "environments" / "my_env"
    legacy.mkdir(parents=True)
    (legacy / em.LEGACY_REL).write_text(json.dumps({"owner": "o", "name": "legacy"}))
    assert em.find_environment_metadata(env_name="my-env", module_name="my_env")["name"] == "legacy"
    target = em.write_environment_metadata(legacy, {"owner": "o", "name": "new"})
    assert target == legacy / em.NEW_REL and not (legacy / em.LEGACY_REL).exists()
    assert em.get_environment_metadata(legacy)["name"] == "new"
    (legacy / em.NEW_REL).write_text("{broken")
    assert em.get_environment_metadata(legacy) is None and em.find_environment_metadata(env_name="nope") is None
    dirs = em.candidate_dirs("my-env", Path("given"), "my_env")
    assert dirs[0] == Path("given") and dirs[-1] == Path(".") and Path("environments") / "my_env" in dirs


# ------------------------------------------------------------------------------------------------ eval push
def test_eval_push_discovers_parses_and_uploads(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    with pytest.raises(FileNotFoundError):
        ep.find_latest_run_dir("gsm8k", "org/m")
    base = tmp_path / "outputs" / "evals" / "gsm8k--org--m"
    old, new = base / "run1", base / "run2"
    old.mkdir(parents=True)
    new.mkdir()
    os.utime(old, (1, 1))
    (new / "metadata.json").write_text(json.dumps({"avg_reward": 0.5, "task_type": "math", "num_examples": 2}))
    (new / "results.jsonl").write_text('{"id": 3, "reward": 1.0, "answer": "4"}\nnot json\n[1]\n\n{"reward": 0.0}\n')
    assert ep.find_latest_run_dir("gsm8k", "org/m").resolve() == new.resolve()
    rows = ep.load_results_jsonl(new / "results.jsonl")
    assert len(rows) == 2 and ep.to_hub_samples(rows)[0] == {"example_id": 3, "reward": 1.0, "answer": "4"} and ep.to_hub_samples(rows)[1]["example_id"] == 0
    assert ep.resolve_upstream("gsm8k", None, "o/gsm8k") == ("o/gsm8k", None) and ep.resolve_upstream("gsm8k", None, None) == (None, None)
    assert ep.push_eval_results_to_hub("gsm8k", "org/m", "job1") is None  # no upstream: nothing uploaded

    import httpx

    from prime_b200.platform.core import Config
    from prime_b200.platform.evals import EvalsClient

    class API:
        base_url, api_key = "https://api.test", "k"

        def __init__(self):
            self.calls, self.config = [], Config(writable=False)

        def request(self, method, endpoint, params=None, json=None, **kw):
            self.calls.append((method, endpoint, json))
            if endpoint == "/evaluations/":
                return {"evaluation_id": "ev9"}
            if endpoint == "/environmentshub/lookup":
                return {"data": {"id": json["id"]}}
            return {}

        def get(self, endpoint, params=None, **kw):
            self.calls.append(("GET", endpoint, None))
            return {"data": {"id": "env-id-1"}}

        def post(self, endpoint, json=None, **kw):
            return self.request("POST", endpoint, json=json)

    uploads = []

    def fake_post(url, content=None, headers=None, timeout=None):
        import json as _json

        uploads.append((url, _json.loads(content), headers))
        return httpx.Response(200, json={}, request=httpx.Request("POST", url))

    monkeypatch.setattr(EvalsClient, "_post", staticmethod(fake_post))
    api = API()
    assert ep.push_eval_results_to_hub("gsm8k", "org/m", "job1", upstream_slug="o/gsm8k", client=api) == "ev9"
    posted = {e: j for m, e, j in api.calls if m == "POST"}
    create = posted["/evaluations/"]
    assert create["environments"] == [{"id": "env-id-1"}] and create["metrics"] == {"avg_reward": 0.5} and create["metadata"]["job_id"] == "job1"
    assert "/evaluations/ev9/finalize" in posted
    (url, body, headers), = uploads
    assert url.endswith("/api/v1/evaluations/ev9/samples") and len(body["samples"]) == 2 and headers["Authorization"] == "Bearer k"


# ------------------------------------------------------------------------------------------------ time helpers
def test_time_helpers():
    now = datetime.now(timezone.utc)
    assert tu.human_age(now - timedelta(seconds=5)) == "5s" and tu.human_age(now - timedelta(hours=3, minutes=10)) == "3h"
    assert tu.human_age((now - timedelta(days=2)).isoformat().replace("+00:00", "Z")) == "2d"
    assert tu.format_time_ago(None) == "-" and tu.format_time_ago(now - timedelta(seconds=10)) == "just now"
    assert tu.format_time_ago(now - timedelta(minutes=5)) == "5m ago"
    old = now - timedelta(days=45)
    assert tu.format_time_ago(old) == old.strftime("%Y-%m-%d")
    assert tu.parse_dt("2026-01-02T03:04:05Z").tzinfo is not None and tu.to_utc(datetime(2026, 1, 1)).tzinfo is timezone.utc
    items = [{"created_at": "2026-01-03T00:00:00Z"}, {"created_at": None}, {"created_at": "2026-01-01T00:00:00Z"}, {"created_at": "garbage"}]
    ordered = tu.sort_by_created(items)
    assert ordered[-1]["created_at"] == "2026-01-03T00:00:00Z" and ordered[-2]["created_at"] == "2026-01-01T00:00:00Z"
    assert tu.iso_timestamp("2026-01-02T03:04:05Z").startswith("2026-01-02")


# ------------------------------------------------------------------------------------------------ hosted eval logs
def test_hosted_eval_log_cleaning_and_overlap():
    raw = "\x1b[32mstart\x1b[0m\n 10%|#         | 1/10 [00:01<00:09]\n100%|##########| 10/10 [00:10<00:00]\nProcessing: 50%\nProcessing: 100% done\nend\n"
    cleaned = he.clean_logs(raw).splitlines()
    # only tqdm bars ("NN%|…|") are progress noise: the 10 % refresh goes, the finished bar and ordinary log lines stay
    assert cleaned == ["start", "100%|##########| 10/10 [00:10<00:00]", "Processing: 50%", "Processing: 100% done", "end"]
    assert he.clean_logs("Waiting for container to start...") == ""
    assert he.get_new_log_lines("", "a\nb") == ["a", "b"]
    assert he.get_new_log_lines("a\nb\nc", "b\nc\nd\ne") == ["d", "e"]  # sliding tail window
    assert he.get_new_log_lines("a\nb", "x\ny") == ["x", "y"] and he.get_new_log_lines("a\nb", "a\nb") == []
    assert he.EvalStatus("COMPLETED") is he.EvalStatus.COMPLETED
    assert he.EvalStatus.TIMEOUT.is_terminal and not he.EvalStatus.RUNNING.is_terminal
    assert he.EvalStatus.terminal_statuses() == {he.EvalStatus.COMPLETED, he.EvalStatus.FAILED, he.EvalStatus.TIMEOUT, he.EvalStatus.CANCELLED}
    assert he.tqdm_line("plain text") is None and he.tqdm_line(" 10%|#   | 1/10") == "" and he.tqdm_line("100%|####| 10/10").startswith("100%")
    body = he.HostedEvalConfig("env-1", "m", 5, 3, env_args={}, timeout_minutes=0, name="n", sampling_args={"t": 1}).payload()
    assert body == {"environment_ids": ["env-1"], "inference_model": "m", "name": "n",
                    "eval_config": {"num_examples": 5, "rollouts_per_example": 3, "timeout_minutes": 0, "allow_sandbox_access": False,
                                    "allow_instances_access": False, "sampling_args": {"t": 1}}}  # fmt: skip


# ------------------------------------------------------------------------------------------------ Connect-RPC helpers
def test_command_session_request_and_event_collection():
    from prime_b200.platform.sandboxes import rpc_command_session as rpc
    from prime_b200.platform.sandboxes import rpc_schema

    req = rpc.build_command_session_start_request("echo hi && false", "/work", {"A": "1"})
    assert req.command.cmd == "/bin/bash" and list(req.command.args) == ["-c", "echo hi && false"] and req.command.cwd == "/work"
    assert dict(req.command.envs) == {"A": "1"} and req.stdin is False
    assert rpc.build_start_request("x", None, None).command.cwd == ""
    # wire round trip: the runtime-built schema serialises like any protobuf message
    again = rpc_schema.StartRequest.FromString(req.SerializeToString())
    assert again.command.args[1] == "echo hi && false"

    def event(**kw):
        r = rpc_schema.StartResponse()
        if "end" in kw:
            r.event.end.exit_code = kw["end"]
        elif "stdout" in kw:
            r.event.data.stdout = kw["stdout"]
        elif "stderr" in kw:
            r.event.data.stderr = kw["stderr"]
        elif "pty" in kw:
            r.event.data.pty = kw["pty"]
        return r

    out, err = [], []
    assert rpc.collect_command_session_start_event(rpc_schema.StartResponse(), out, err) is None
    assert rpc.collect_command_session_start_event(event(stdout=b"hi\n"), out, err) is None
    assert rpc.collect_command_session_start_event(event(stderr=b"warn \xff"), out, err) is None
    assert rpc.collect_command_session_start_event(event(pty=b"tty"), out, err) is None
    assert rpc.collect_command_session_start_event(event(end=3), out, err) == 3
    assert out == ["hi\n", "tty"] and err == ["warn �"]
    c = rpc.OutputCollector()
    for e in (event(stdout=b"a"), event(stdout=b"b"), event(end=0)):
        c.feed(e)
    assert c.result() == ("ab", "", 0)
