"""Root CLI: panels, --version, --context, update banner; plus smoke flows of the account/compute groups against a fake API
(reference tests: packages/prime/tests/test_main*.py, test_config_*.py, test_pods_*.py, test_secrets*.py)."""

import json

import pytest
from typer.testing import CliRunner

from prime_b200.platform import __version__
from prime_b200.platform import main as main_mod
from prime_b200.platform.commands import availability as av_mod
from prime_b200.platform.commands import disks as disks_mod
from prime_b200.platform.commands import pods as pods_mod
from prime_b200.platform.commands import secrets as secrets_mod
from prime_b200.platform.commands import teams as teams_mod
from prime_b200.platform.commands import whoami as whoami_mod
from prime_b200.platform.core import Config
from prime_b200.platform.main import GROUPS, app

runner = CliRunner()


def test_root_help_lists_every_group_in_its_panel():
    out = runner.invoke(app, ["--help"]).output
    for name, _, panel in GROUPS:
        assert name in out and panel in out
    assert len(GROUPS) == 21
    assert runner.invoke(app, ["--version"]).output.strip() == f"Prime CLI version: {__version__}"


def test_every_command_renders_help():
    import click
    import typer

    failures = []

    def walk(cmd, path):
        r = runner.invoke(app, path + ["--help"])
        if r.exit_code != 0:
            failures.append((path, r.output[-200:]))
        if isinstance(cmd, click.Group):
            for name, sub in cmd.commands.items():
                walk(sub, path + [name])

    walk(typer.main.get_command(app), [])
    assert not failures, failures


def test_context_flag_sets_env_for_subcommand(isolated_home, monkeypatch):
    c = Config()
    c.set_api_key("k-prod")
    c.set_base_url("https://staging.example")
    c.set_api_key("k-staging")
    c.save_environment("staging")
    c.load_environment("production")
    r = runner.invoke(app, ["--context", "ghost", "config", "view"])
    assert r.exit_code == 1 and "Unknown context 'ghost'" in r.output and "staging" in r.output
    monkeypatch.delenv("PRIME_CONTEXT", raising=False)
    r = runner.invoke(app, ["--context", "staging", "config", "view"])
    assert r.exit_code == 0 and "staging.example" in r.output
    monkeypatch.delenv("PRIME_CONTEXT", raising=False)
    assert "staging.example" not in runner.invoke(app, ["config", "view"]).output


def test_update_banner_only_with_subcommand(monkeypatch):
    monkeypatch.setattr(main_mod, "check_for_update", lambda v=None: (True, "9.9.9"))
    r = runner.invoke(app, ["config", "envs"])
    assert "new version of prime is available: 9.9.9" in r.output
    assert "9.9.9" not in runner.invoke(app, ["--version"]).output


def test_whoami_and_teams(fake_api, isolated_home):
    fake_api({("GET", "/user/whoami"): {"data": {"id": "u1", "email": "a@b.c", "name": "Ann", "scope": {"pods": {"read": True, "write": False}}}},
              ("GET", "/user/teams"): {"data": [{"teamId": "t1", "name": "Core", "slug": "core", "role": "admin", "createdAt": "2026"}], "total_count": 1},
              ("GET", "/teams/t1/members"): {"data": [{"userId": "u1", "userName": "Ann", "userEmail": "a@b.c", "role": "admin", "joinedAt": "2026"}]}},
             whoami_mod, teams_mod)  # fmt: skip
    r = runner.invoke(app, ["whoami"])
    assert r.exit_code == 0 and "a@b.c" in r.output and Config().user_id == "u1"
    out = json.loads(runner.invoke(app, ["teams", "list", "-o", "json"]).output)
    assert out["teams"][0]["slug"] == "core"
    assert "Ann" in runner.invoke(app, ["teams", "members", "--team-id", "t1"]).output


def test_global_secrets_are_team_scoped(fake_api, isolated_home):
    Config().set_team("t1", "Core", "admin")
    api = fake_api({("GET", "/secrets/"): {"data": [{"id": "s1", "name": "HF_TOKEN", "updatedAt": "2026-01-02T03:04:05Z"}]},
                    ("POST", "/secrets/"): lambda params=None, json=None: {"data": {"id": "s2", **json}},
                    ("GET", "/secrets/s1"): {"data": {"id": "s1", "name": "HF_TOKEN"}}, ("DELETE", "/secrets/s1"): {}}, secrets_mod)  # fmt: skip
    r = runner.invoke(app, ["secret", "list"])
    assert r.exit_code == 0 and "HF_TOKEN" in r.output and api.calls[0][2] == {"teamId": "t1"}
    r = runner.invoke(app, ["secret", "create", "--name", "WANDB_KEY", "--value", "v"])
    assert r.exit_code == 0, r.output
    assert api.called("POST", "/secrets/")[0][3]["teamId"] == "t1"
    assert runner.invoke(app, ["secret", "delete", "s1", "-y"]).exit_code == 0 and api.called("DELETE", "/secrets/s1")[0][2] == {"teamId": "t1"}


POD = {"id": "p1", "name": "box", "gpuName": "B200_180GB", "gpuCount": 8, "status": "ACTIVE", "createdAt": "2026-01-02T03:04:05Z",
       "providerType": "hyperstack", "priceHr": 31.2, "sshConnection": "root@1.2.3.4 -p 2222", "ip": "1.2.3.4"}  # fmt: skip


def test_pods_list_status_terminate(fake_api):
    api = fake_api({("GET", "/pods"): {"total_count": 1, "offset": 0, "limit": 100, "data": [POD]},
                    ("GET", "/pods/p1"): POD, ("DELETE", "/pods/p1"): {},
                    ("GET", "/pods/status"): {"data": [{"podId": "p1", "providerType": "hyperstack", "status": "ACTIVE", "sshConnection": "root@1.2.3.4 -p 2222", "ip": "1.2.3.4"}]}},
                   pods_mod)  # fmt: skip
    r = runner.invoke(app, ["pods", "list"])
    assert r.exit_code == 0 and "box" in r.output and "B200_180GB" in r.output
    out = json.loads(runner.invoke(app, ["pods", "list", "-o", "json"]).output)
    assert out["total_count"] == 1 and out["pods"][0]["id"] == "p1"
    r = runner.invoke(app, ["pods", "status", "p1", "-o", "json"])
    assert r.exit_code == 0 and json.loads(r.output)["status"] == "ACTIVE"
    r = runner.invoke(app, ["pods", "terminate", "p1"], input="n\n")
    assert not api.called("DELETE", "/pods/p1")
    assert runner.invoke(app, ["pods", "terminate", "p1", "-y"]).exit_code == 0 and api.called("DELETE", "/pods/p1")
    assert pods_mod.split_ssh_target("root@1.2.3.4 -p 2222") == ("root@1.2.3.4", "2222")
    assert pods_mod.ssh_command("/k", "root@h -p 22")[:3] == ["ssh", "-i", "/k"]


def test_disks_and_gpu_types(fake_api):
    disk = {"id": "d1", "name": "data", "size": 500, "status": "ACTIVE", "createdAt": "2026-01-02T03:04:05Z", "updatedAt": "2026-01-02T03:04:05Z",
            "providerType": "hyperstack", "priceHr": 0.05}  # fmt: skip
    fake_api({("GET", "/disks"): {"total_count": 1, "offset": 0, "limit": 100, "data": [disk]}, ("PATCH", "/disks/d1"): {"ok": True},
              ("GET", "/availability/gpu-summary"): {"H100_80GB": {}, "B200_180GB": {}}}, disks_mod, av_mod)  # fmt: skip
    r = runner.invoke(app, ["disks", "list"])
    assert r.exit_code == 0 and "data" in r.output, r.output
    assert runner.invoke(app, ["disks", "update", "d1", "--name", "new"]).exit_code == 0
    r = runner.invoke(app, ["availability", "gpu-types"])
    assert r.exit_code == 0 and r.output.index("B200_180GB") < r.output.index("H100_80GB")


@pytest.mark.parametrize("argv", [["pods", "list"], ["secret", "list"], ["teams", "list"]])
def test_api_errors_become_one_line_exit_1(argv, isolated_home):
    r = runner.invoke(app, argv)  # no API key configured → typed auth error, not a traceback
    assert r.exit_code == 1 and "Error" in r.output and "Traceback" not in r.output
