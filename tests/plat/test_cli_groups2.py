"""Third batch: deployments state machine, images push/delete references, config commands, switch, inference models,
availability tables (reference tests: packages/prime/tests/test_deployments.py, test_images.py, test_config_cmd.py,
test_switch.py, test_availability.py)."""

import io
import json
import tarfile

import httpx
import pytest
from typer.testing import CliRunner

from prime_b200.platform.api.deployments import Adapter
from prime_b200.platform.commands import availability as av_mod
from prime_b200.platform.commands import config as cfg_mod
from prime_b200.platform.commands import deployments as dep_mod
from prime_b200.platform.commands import images as img_mod
from prime_b200.platform.commands import switch as sw_mod
from prime_b200.platform.core import Config
from prime_b200.platform.main import app

runner = CliRunner()
T = "2026-01-02T03:04:05Z"


def adapter(status="READY", dep="NOT_DEPLOYED", **kw):
    return {"id": "a1", "displayName": "tuned", "userId": "u", "rftRunId": "r", "baseModel": "Qwen/Qwen3-4B", "status": status, "deploymentStatus": dep,
            "createdAt": T, "updatedAt": T, **kw}  # fmt: skip


def test_deployment_state_machine(fake_api):
    A = lambda **kw: Adapter.model_validate(adapter(**kw))  # noqa: E731
    assert dep_mod.deploy_blocker(A()) is None
    assert dep_mod.deploy_blocker(A(status="UPLOADING"))[1] == 1
    assert dep_mod.deploy_blocker(A(dep="DEPLOYED")) == ("Model is already deployed.", 0)  # idempotent: exit 0
    assert dep_mod.deploy_blocker(A(dep="DEPLOYING"))[1] == 1 and dep_mod.unload_blocker(A(dep="UNLOADING"))[1] == 1
    assert dep_mod.unload_blocker(A(dep="NOT_DEPLOYED")) == ("Model is not deployed.", 0) and dep_mod.unload_blocker(A(dep="DEPLOYED")) is None
    api = fake_api({("GET", "/rft/adapters"): {"adapters": [adapter()], "total": 1}, ("GET", "/rft/adapters/a1"): {"adapter": adapter()},
                    ("GET", "/rft/deployable-models"): {"models": ["Qwen/Qwen3-4B"]},
                    ("POST", "/rft/adapters/a1/deploy"): {"adapter": adapter(dep="DEPLOYING")}}, dep_mod)  # fmt: skip
    assert "tuned" in runner.invoke(app, ["deployments", "list"]).output
    r = runner.invoke(app, ["deployments", "create", "a1"], input="n\n")
    assert "Cancelled" in r.output and not api.called("POST", "/rft/adapters/a1/deploy")
    r = runner.invoke(app, ["deployments", "create", "a1", "-y"])
    assert r.exit_code == 0 and "DEPLOYING" in r.output and "Qwen/Qwen3-4B:a1" in r.output
    api.routes[("GET", "/rft/deployable-models")] = {"models": []}
    assert runner.invoke(app, ["deployments", "create", "a1", "-y"]).exit_code == 1
    r = runner.invoke(app, ["deployments", "delete", "a1"])
    assert r.exit_code == 0 and "not deployed" in r.output


def test_image_references_and_context_packaging(tmp_path):
    assert img_mod.split_reference("myapp") == ("myapp", "latest") and img_mod.split_reference("myapp:v1") == ("myapp", "v1")
    with pytest.raises(ValueError):
        img_mod.split_reference("myapp", require_tag=True)
    assert img_mod.parse_delete_reference("team-t9/app:v1", None) == ("app", "v1", "t9")
    assert img_mod.parse_delete_reference("app:v1", "tdef") == ("app", "v1", "tdef")
    for bad in ("library/app:v1", "team-/app:v1", "app"):
        with pytest.raises(ValueError):
            img_mod.parse_delete_reference(bad, None)
    ctx = tmp_path / "ctx"
    ctx.mkdir()
    (ctx / "main.py").write_text("print(1)")
    df = tmp_path / "Custom.Dockerfile"
    df.write_text("FROM python:3.12-slim\n")
    size = img_mod.package_context(ctx, df, tmp_path / "c.tgz")
    names = tarfile.open(tmp_path / "c.tgz").getnames()
    assert size > 0 and "./main.py" in names and img_mod.PACKAGED_DOCKERFILE_PATH in names


def test_image_push_flow(tmp_path, fake_api, monkeypatch, isolated_home):
    Config().set_team("t1", "Core", "admin")
    ctx = tmp_path / "ctx"
    ctx.mkdir()
    (ctx / "Dockerfile").write_text("FROM scratch\n")
    puts = []
    monkeypatch.setattr(img_mod.httpx, "put", lambda url, content=None, **kw: (puts.append((url, len(content.read()))), httpx.Response(200, request=httpx.Request("PUT", url)))[1])
    api = fake_api({("POST", "/images/build"): {"build_id": "b1", "upload_url": "https://s3/ctx", "fullImagePath": "reg/team-t1/app:v2"},
                    ("POST", "/images/build/b1/start"): {}}, img_mod)  # fmt: skip
    r = runner.invoke(app, ["images", "push", "app:v2", "-c", str(ctx), "--platform", "linux/arm64"])
    assert r.exit_code == 0, r.output
    assert api.calls[0][3] == {"image_name": "app", "image_tag": "v2", "dockerfile_path": img_mod.PACKAGED_DOCKERFILE_PATH, "platform": "linux/arm64", "team_id": "t1"}
    assert puts and puts[0][0] == "https://s3/ctx" and puts[0][1] > 0 and api.calls[1][3] == {"context_uploaded": True}
    assert "reg/team-t1/app:v2" in r.output
    assert runner.invoke(app, ["images", "push", "ns/app:v2", "-c", str(ctx)]).exit_code == 1
    assert "Dockerfile not found" in runner.invoke(app, ["images", "push", "app", "-c", str(tmp_path)]).output


def test_config_commands_roundtrip(isolated_home, monkeypatch):
    monkeypatch.setattr(cfg_mod, "_remember_user", lambda config, key: None)  # no network
    assert runner.invoke(app, ["config", "set-api-key", "pit_0123456789abcdef"]).exit_code == 0
    out = runner.invoke(app, ["config", "view"]).output
    assert "pit_0123456789abcdef" not in out and "cdef" in out  # masked
    assert runner.invoke(app, ["config", "set-base-url", "https://api.staging.example/"]).exit_code == 0
    assert Config().base_url == "https://api.staging.example"
    assert runner.invoke(app, ["config", "set-ssh-key-path", "~/.ssh/id_ed25519"]).exit_code == 0
    assert runner.invoke(app, ["config", "set-share-resources-with-team", "true"]).exit_code == 0 and Config().share_resources_with_team
    assert runner.invoke(app, ["config", "save", "Staging Env!"]).exit_code == 0
    envs = runner.invoke(app, ["config", "envs"]).output
    assert "production" in envs and "staging" in envs.lower()
    assert runner.invoke(app, ["config", "use", "production"]).exit_code == 0 and "staging" not in Config().base_url
    assert runner.invoke(app, ["config", "use", "nope"]).exit_code == 1
    assert runner.invoke(app, ["config", "set-team-id", "not a valid id!"]).exit_code == 1
    r = runner.invoke(app, ["config", "reset"], input="n\n")
    assert Config().api_key
    assert runner.invoke(app, ["config", "reset", "-y"]).exit_code == 0 and not Config().api_key
    # env var wins over the file and is reported as such
    monkeypatch.setenv("PRIME_API_KEY", "pit_from_env_000000")
    assert "env" in runner.invoke(app, ["config", "view"]).output.lower()


TEAMS = [{"teamId": "t1", "name": "Core", "slug": "core", "role": "ADMIN"}, {"teamId": "t2", "name": "Research", "slug": "research", "role": "member"}]


def test_switch_personal_slug_id_interactive(isolated_home, monkeypatch):
    monkeypatch.setattr(sw_mod, "fetch_teams", lambda client: TEAMS)
    Config().set_api_key("k")
    assert runner.invoke(app, ["switch", "research"]).exit_code == 0 and Config().team_id == "t2" and Config().team_name == "Research"
    assert runner.invoke(app, ["switch", "t1"]).exit_code == 0 and Config().team_id == "t1"
    assert runner.invoke(app, ["switch", "personal"]).exit_code == 0 and Config().team_id is None
    r = runner.invoke(app, ["switch", "ghost"])
    assert r.exit_code == 1 and "core" in r.output and "research" in r.output  # suggests valid slugs
    r = runner.invoke(app, ["switch"], input="2\n")
    assert r.exit_code == 0 and Config().team_id in ("t1", "t2")


def test_inference_models_table(monkeypatch):
    from prime_b200.platform.commands import inference as inf_mod

    class Fake:
        def list_models(self):
            return {"data": [{"id": "meta/llama", "created": 1767322245, "pricing": {"input_usd_per_mtok": 0.2, "output_usd_per_mtok": 0.6}}, {"id": "q/qwen"}]}

    monkeypatch.setattr(inf_mod, "InferenceClient", Fake)
    r = runner.invoke(app, ["inference", "models"])
    assert r.exit_code == 0 and "meta/llama" in r.output and "q/qwen" in r.output
    out = json.loads(runner.invoke(app, ["inference", "models", "-o", "json"]).output)
    assert [m["id"] for m in out["data"]] == ["meta/llama", "q/qwen"]  # raw OpenAI-style payload


def test_availability_list_groups_and_short_ids(fake_api):
    offer = {"cloudId": "c1", "gpuType": "B200_180GB", "socket": "SXM6", "provider": "hyperstack", "dataCenter": "dc1", "country": "US", "gpuCount": 8,
             "gpuMemory": 180, "security": "secure_cloud", "prices": {"onDemand": 31.2, "currency": "USD"}, "stockStatus": "Available",
             "vcpu": {"defaultCount": 128}, "memory": {"defaultCount": 1024}, "disk": {"defaultCount": 2000}, "isSpot": False, "images": ["ubuntu_22_cuda_12"]}  # fmt: skip
    api = fake_api({("GET", "/availability/gpus"): {"items": [offer, {**offer, "cloudId": "c2", "gpuCount": 1, "prices": {"onDemand": 4.1, "currency": "USD"}}], "totalCount": 2}}, av_mod)
    # /availability/multi-node is not routed → that endpoint "is down": the note goes to stderr, stdout stays valid JSON
    r = runner.invoke(app, ["availability", "list", "--gpu-type", "B200_180GB"])
    assert r.exit_code == 0, r.output
    assert "B200 180GB" in r.output and "hyperstack" in r.output  # display form: underscores → spaces
    res = CliRunner(mix_stderr=False).invoke(app, ["availability", "list", "--gpu-type", "B200_180GB", "-o", "json"])
    assert "multi-node" in res.stderr
    out = json.loads(res.stdout)
    rows = out.get("gpu_resources") or out.get("gpus") or next(iter(out.values()))
    ids = [x["id"] for x in rows]
    assert len(set(ids)) == 2 and all(len(i) == 6 for i in ids)  # md5[:6] short ids, stable per offer
    assert api.calls[0][2].get("gpu_type") == "B200_180GB"
