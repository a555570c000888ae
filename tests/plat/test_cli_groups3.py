"""Fourth batch: pods (list/status/history/terminate + the pure helpers of create/ssh), disks, secrets, teams, registry, whoami
(reference tests: packages/prime/tests/test_pods.py, test_disks.py, test_secrets.py, test_teams.py, test_whoami.py)."""

import json

from typer.testing import CliRunner

from prime_b200.platform.commands import disks as disks_mod
from prime_b200.platform.commands import pods as pods_mod
from prime_b200.platform.commands import registry as reg_mod
from prime_b200.platform.commands import secrets as sec_mod
from prime_b200.platform.commands import teams as teams_mod
from prime_b200.platform.commands import whoami as who_mod
from prime_b200.platform.core import Config
from prime_b200.platform.main import app

runner = CliRunner()
T = "2026-01-02T03:04:05Z"


def pod(pid="p1", **kw):
    return {"id": pid, "name": "trainer", "gpuName": "B200_180GB", "gpuCount": 8, "status": "ACTIVE", "createdAt": T, "providerType": "dc1",
            "priceHr": 31.6, "teamId": None, **kw}  # fmt: skip


def test_pods_list_status_history_terminate(fake_api):
    status = {"podId": "p1", "providerType": "dc1", "status": "ACTIVE", "sshConnection": ["root@1.2.3.4 -p 2222", None], "ip": ["1.2.3.4"],
              "priceHr": 31.6, "installationProgress": 80,
              "primePortMapping": [{"internal": "8888", "external": "31000", "protocol": "TCP", "description": "jupyter"}]}  # fmt: skip
    hist = {"id": "p0", "name": "old", "providerType": "dc1", "type": "HOSTED", "createdAt": T, "terminatedAt": T, "gpuName": "H100_80GB", "gpuCount": 1,
            "priceHr": 2.0, "userId": "u1", "totalBilledPrice": 12.5}  # fmt: skip
    api = fake_api({("GET", "/pods"): {"total_count": 1, "offset": 0, "limit": 100, "data": [pod()]},
                    ("GET", "/pods/status"): {"data": [status]}, ("GET", "/pods/p1"): pod(),
                    ("GET", "/pods/history"): {"total_count": 1, "offset": 0, "limit": 100, "data": [hist]},
                    ("DELETE", "/pods/p1"): {}}, pods_mod)  # fmt: skip
    r = runner.invoke(app, ["pods", "list"])
    assert r.exit_code == 0 and "trainer" in r.output and "B200_180GB" in r.output and "31.60" in r.output
    data = json.loads(runner.invoke(app, ["pods", "list", "--output", "json"]).output)
    assert data["total_count"] == 1 and data["pods"][0]["id"] == "p1"
    assert runner.invoke(app, ["pods", "list", "--watch", "--output", "json"]).exit_code == 1  # incompatible flags
    r = runner.invoke(app, ["pods", "status", "p1"])
    assert r.exit_code == 0 and "root@1.2.3.4 -p 2222" in r.output and "80%" in r.output and "31000" in r.output
    assert api.called("GET", "/pods/status")[0][2] == {"pod_ids": ["p1"]}
    r = runner.invoke(app, ["pods", "history"])
    assert r.exit_code == 0 and "$12.50" in r.output and "H100_80GB" in r.output
    r = runner.invoke(app, ["pods", "terminate", "p1"], input="n\n")
    assert "cancelled" in r.output.lower() and not api.called("DELETE", "/pods/p1")
    r = runner.invoke(app, ["pods", "terminate", "p1", "-y"])
    assert r.exit_code == 0 and api.called("DELETE", "/pods/p1")


def test_pod_helpers():
    assert pods_mod.split_ssh_target("root@1.2.3.4 -p 2222") == ("root@1.2.3.4", "2222")
    assert pods_mod.split_ssh_target("ubuntu@host") == ("ubuntu@host", "22")
    cmd = pods_mod.ssh_command("/k/id", "root@1.2.3.4 -p 2222")
    assert cmd[:3] == ["ssh", "-i", "/k/id"] and cmd[-3:] == ["-p", "2222", "root@1.2.3.4"]
    assert pods_mod.parse_env_pairs(["A=1", "B=x=y"]) == [{"key": "A", "value": "1"}, {"key": "B", "value": "x=y"}]
    assert pods_mod.valid_pod_name("my-pod-1") and not pods_mod.valid_pod_name("bad name!") and not pods_mod.valid_pod_name("1234")


def test_pod_status_without_data_is_an_error(fake_api):
    fake_api({("GET", "/pods/status"): {"data": []}}, pods_mod)
    r = runner.invoke(app, ["pods", "status", "nope"])
    assert r.exit_code == 1 and "No status found" in r.output


def disk(**kw):
    return {"id": "d1", "name": "data", "createdAt": T, "updatedAt": T, "status": "ACTIVE", "providerType": "dc1", "size": 500, "priceHr": 0.07,
            "info": {"country": "US", "dataCenterId": "dc-a"}, "pods": ["p1"], "clusters": [], **kw}  # fmt: skip


def test_disks_list_get_update_terminate(fake_api):
    api = fake_api({("GET", "/disks"): {"total_count": 1, "offset": 0, "limit": 100, "data": [disk()]}, ("GET", "/disks/d1"): disk(),
                    ("PATCH", "/disks/d1"): {"status": "ok"}, ("DELETE", "/disks/d1"): {"status": "deleted"}}, disks_mod)  # fmt: skip
    r = runner.invoke(app, ["disks", "list"])
    assert r.exit_code == 0 and "data" in r.output and "500" in r.output
    got = json.loads(runner.invoke(app, ["disks", "get", "d1", "--output", "json"]).output)
    assert got["id"] == "d1" and got["pods"] == ["p1"]
    assert runner.invoke(app, ["disks", "update", "d1", "--name", "renamed"]).exit_code == 0
    assert api.called("PATCH", "/disks/d1")[0][3] == {"name": "renamed"}
    r = runner.invoke(app, ["disks", "terminate", "d1"], input="n\n")
    assert not api.called("DELETE", "/disks/d1")
    assert runner.invoke(app, ["disks", "terminate", "d1", "-y"]).exit_code == 0 and api.called("DELETE", "/disks/d1")
    cfg = disks_mod.build_disk_config(100, "scratch", "team9", country="DE")
    assert cfg["disk"]["size"] == 100 and cfg["disk"]["name"] == "scratch" and cfg["disk"]["country"] == "DE" and cfg["team"] == {"teamId": "team9"}


def test_secrets_crud_and_team_scope(fake_api):
    sec = {"id": "s1", "name": "HF_TOKEN", "description": "hub", "isFile": False, "createdAt": T, "updatedAt": T}
    api = fake_api({("GET", "/secrets/"): {"data": [sec]}, ("POST", "/secrets/"): {"data": sec}, ("GET", "/secrets/s1"): {"data": sec},
                    ("PATCH", "/secrets/s1"): {"data": {**sec, "description": "new"}}, ("DELETE", "/secrets/s1"): {}}, sec_mod)  # fmt: skip
    assert "HF_TOKEN" in runner.invoke(app, ["secret", "list"]).output
    assert api.called("GET", "/secrets/")[0][2] is None  # personal scope: no teamId filter
    r = runner.invoke(app, ["secret", "create", "--name", "HF_TOKEN", "--value", "hf_x", "--description", "hub"])
    assert r.exit_code == 0 and "Created" in r.output
    body = api.called("POST", "/secrets/")[0][3]
    assert body["name"] == "HF_TOKEN" and body["value"] == "hf_x" and "teamId" not in body
    assert runner.invoke(app, ["secret", "create", "--name", "not valid!", "--value", "x"]).exit_code == 1
    assert runner.invoke(app, ["secret", "update", "s1", "--description", "new"]).exit_code == 0
    assert api.called("PATCH", "/secrets/s1")[0][3] == {"description": "new"}
    got = json.loads(runner.invoke(app, ["secret", "get", "s1", "--output", "json"]).output)
    assert got["name"] == "HF_TOKEN" and "value" not in got
    assert runner.invoke(app, ["secret", "delete", "s1", "-y"]).exit_code == 0 and api.called("DELETE", "/secrets/s1")
    c = Config()
    c.set_team("team9", team_name="Lab", team_role="admin")
    api.calls.clear()
    runner.invoke(app, ["secret", "list"])
    assert api.called("GET", "/secrets/")[0][2] == {"teamId": "team9"}


def test_teams_list_paginates_and_members_need_a_team(fake_api):
    teams = [{"teamId": f"t{i}", "name": f"Team {i}", "slug": f"team-{i}", "role": "member", "createdAt": T} for i in range(3)]

    def page(params=None, json=None):
        o, n = params["offset"], params["limit"]
        return {"data": teams[o : o + n], "total_count": len(teams)}

    api = fake_api({("GET", "/user/teams"): page, ("GET", "/teams/t1/members"): {"data": [{"userId": "u1", "userName": "Ada", "role": "admin", "joinedAt": T}]}},
                   teams_mod)  # fmt: skip
    assert [t["teamId"] for t in teams_mod.fetch_teams(api, page=2)] == ["t0", "t1", "t2"] and len(api.called("GET", "/user/teams")) == 2
    r = runner.invoke(app, ["teams", "list"])
    assert r.exit_code == 0 and "team-2" in r.output
    r = runner.invoke(app, ["teams", "members"])
    assert r.exit_code == 1 and "No team selected" in r.output
    r = runner.invoke(app, ["teams", "members", "--team-id", "t1"])
    assert r.exit_code == 0 and "Ada" in r.output


def test_registry_list_and_check_image(fake_api):
    cred = {"id": "c1", "name": "ghcr", "server": "ghcr.io", "createdAt": T, "updatedAt": T, "userId": "u", "teamId": None}
    api = fake_api({("GET", "/template/registry-credentials"): {"credentials": [cred]},
                    ("POST", "/template/check-docker-image"): lambda params=None, json=None: {"accessible": json["image"].startswith("ghcr.io/ok"), "details": "pull ok" if json["image"].startswith("ghcr.io/ok") else "denied"}},
                   reg_mod)  # fmt: skip
    r = runner.invoke(app, ["registry", "list"])
    assert r.exit_code == 0 and "ghcr.io" in r.output and "user:u" in r.output  # scope: team id, else user:<id>, else personal
    assert runner.invoke(app, ["registry", "check-image", "ghcr.io/ok/app:1"]).exit_code == 0
    r = runner.invoke(app, ["registry", "check-image", "ghcr.io/private/app:1", "--registry-credentials-id", "c1"])
    assert r.exit_code == 1 and "denied" in r.output
    assert api.called("POST", "/template/check-docker-image")[-1][3]["registry_credentials_id"] == "c1"


def test_whoami_persists_user_id_and_shows_scopes(fake_api):
    fake_api({("GET", "/user/whoami"): {"data": {"id": "u-42", "slug": "ada", "name": "Ada L", "email": "ada@example.com",
                                                  "scope": {"pods": {"read": True, "write": False}, "billing": None}}}}, who_mod)  # fmt: skip
    r = runner.invoke(app, ["whoami"])
    assert r.exit_code == 0 and "u-42" in r.output and "ada@example.com" in r.output and "Personal" in r.output and "pods" in r.output
    assert Config(writable=False).user_id == "u-42"
    doc = json.loads(runner.invoke(app, ["whoami", "-o", "json"]).output)
    assert doc["user"]["slug"] == "ada" and doc["team"] is None and doc["scope"]["billing"] is None
    assert who_mod.permission_rows({"pods": {"read": True, "write": False}, "billing": None}) == [("pods", "✓", "✗"), ("billing", "-", "-")]
    fake_api({("GET", "/user/whoami"): {"data": "nope"}}, who_mod)
    assert runner.invoke(app, ["whoami"]).exit_code == 1


def test_reference_pagination_flags_on_rl_and_deployments(fake_api):
    """`--team/-t`, `--num/-n`, `--page/-p` as in the reference (rl list pages client-side: the endpoint returns everything)."""
    from prime_b200.platform.commands import deployments as dep_mod
    from prime_b200.platform.commands import rl as rl_mod

    def run(i):
        return {"id": f"r{i}", "name": f"run{i}", "userId": "u", "teamId": None, "status": "RUNNING", "baseModel": "Qwen/Qwen3-4B",
                "environments": [{"id": "gsm8k"}], "rolloutsPerExample": 8, "seqLen": 2048, "maxSteps": 10, "batchSize": 32,
                "createdAt": f"2026-01-{i + 1:02d}T00:00:00Z", "updatedAt": T}  # fmt: skip

    api = fake_api({("GET", "/rft/runs"): {"runs": [run(i) for i in range(5)]}}, rl_mod)
    out = json.loads(runner.invoke(app, ["rl", "list", "-n", "2", "-p", "2", "-t", "team9", "--output", "json"]).output)
    assert [r["id"] for r in out["runs"]] == ["r2", "r1"] and out["total"] == 5 and out["page"] == 2 and out["per_page"] == 2
    assert api.called("GET", "/rft/runs")[0][2] == {"team_id": "team9"}
    assert "No more results" in runner.invoke(app, ["rl", "ls", "--page", "9"]).output
    assert runner.invoke(app, ["rl", "list", "--num", "0"]).exit_code == 1
    api2 = fake_api({("GET", "/rft/adapters"): {"adapters": [], "total": 45}, ("GET", "/rft/deployable-models"): {"models": ["Qwen/Qwen3-4B"]}}, dep_mod)
    out = json.loads(runner.invoke(app, ["deployments", "list", "-n", "20", "-p", "3", "-t", "team9", "--output", "json"]).output)
    assert out["total"] == 45 and out["page"] == 3 and out["per_page"] == 20
    assert api2.called("GET", "/rft/deployable-models")  # every row is marked deployable / not deployable, as in the reference
    assert api2.called("GET", "/rft/adapters")[0][2] == {"team_id": "team9", "limit": 20, "offset": 40}
    for argv in (["eval", "list", "--help"], ["eval", "push", "--help"]):
        helptext = runner.invoke(app, argv).output
        assert "--env" in helptext
