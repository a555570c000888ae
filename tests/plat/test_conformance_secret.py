"""Behavioural conformance of ``prime secret …`` (account-level secrets) with the reference CLI, personal and team scope
(scenarios: packages/prime/tests/test_secrets.py:109-540; harness and fake hub are ours)."""

import json

import pytest
from typer.testing import CliRunner

from prime_b200.platform.core import client as core_client
from prime_b200.platform.main import app

runner = CliRunner()
SID = "secret-id-1234567890"


def secret(sid, name, description=None, team=None):
    return {"id": sid, "name": name, "description": description, "isFile": False, "userId": None if team else "user-123", "teamId": team,
            "createdAt": "2026-01-15T10:00:00Z", "updatedAt": "2026-01-15T10:00:00Z"}  # fmt: skip


PERSONAL = [secret(SID, "MY_SECRET", "Test secret"), secret("secret-id-0987654321", "API_KEY")]


class Server:
    def __init__(self, secrets):
        self.secrets, self.writes, self.list_params = list(secrets), [], []

    def get(self, endpoint, params=None, **kw):
        if endpoint.rstrip("/") == "/secrets":
            self.list_params.append(params)
            return {"data": self.secrets, "totalCount": len(self.secrets)}
        tail = endpoint.rsplit("/", 1)[1]
        return {"data": next((s for s in self.secrets if s["id"].startswith(tail)), self.secrets[0] if self.secrets else {})}

    def post(self, endpoint, json=None, **kw):
        self.writes.append(("POST", endpoint, json))
        return {"data": {"id": "new-secret-id-001", "name": (json or {}).get("name"), "description": (json or {}).get("description"),
                         "isFile": (json or {}).get("isFile", False)}}  # fmt: skip

    def patch(self, endpoint, json=None, **kw):
        self.writes.append(("PATCH", endpoint, json))
        return {"data": {**self.secrets[0], **{k: v for k, v in (json or {}).items() if k != "value"}}}

    def delete(self, endpoint, **kw):
        self.writes.append(("DELETE", endpoint, None))
        return {}


@pytest.fixture
def hub(monkeypatch):
    def install(secrets=PERSONAL, team=None):
        srv = Server(secrets)
        monkeypatch.setenv("PRIME_API_KEY", "test-key")
        if team:
            monkeypatch.setenv("PRIME_TEAM_ID", team)
        for verb in ("get", "post", "patch", "delete"):
            monkeypatch.setattr(core_client.APIClient, verb, lambda self, *a, _v=verb, **kw: getattr(srv, _v)(*a, **kw))
        return srv

    return install


def run(*argv, input=None):
    return runner.invoke(app, ["secret", *argv], input=input)


def test_list_personal(hub):
    hub()
    r = run("list")
    assert r.exit_code == 0 and all(x in r.output for x in ("Personal Secrets", "MY_SECRET", "API_KEY", "Test secret"))
    doc = json.loads(run("list", "-o", "json").output)
    assert len(doc["secrets"]) == 2 and doc["secrets"][0]["name"] == "MY_SECRET"


def test_list_empty_personal_and_team(hub):
    hub(secrets=[])
    r = run("list")
    assert r.exit_code == 0 and "No personal secrets found" in r.output
    hub(secrets=[], team="team-123")
    r = run("list")
    assert r.exit_code == 0 and "No team secrets found" in r.output


def test_list_team_scope(hub):
    srv = hub(secrets=[secret("team-secret-001", "TEAM_DB_URL", "Shared database URL", team="team-123")], team="team-123")
    r = run("list")
    assert r.exit_code == 0 and "Team Secrets" in r.output and "TEAM_DB_URL" in r.output
    assert (srv.list_params[-1] or {}).get("teamId") == "team-123"  # the listing is asked FOR the team
    doc = json.loads(run("list", "-o", "json").output)
    assert len(doc["secrets"]) == 1 and doc["secrets"][0]["name"] == "TEAM_DB_URL"


def test_create(hub):
    srv = hub()
    r = run("create", "-n", "NEW_SECRET", "-v", "secret-value")
    assert r.exit_code == 0 and "Created personal secret 'NEW_SECRET'" in r.output and "ID:" in r.output
    r = run("create", "-n", "NEW_SECRET", "-v", "secret-value", "-d", "A new test secret")
    assert r.exit_code == 0 and srv.writes[-1][2]["description"] == "A new test secret"
    doc = json.loads(run("create", "-n", "NEW_SECRET", "-v", "value", "-o", "json").output)
    assert doc["name"] == "NEW_SECRET" and "id" in doc
    r = run("create", input="MY_NEW_SECRET\nsecret-value\n")
    assert r.exit_code == 0 and "Created personal secret" in r.output
    r = run("create", "-n", "FILE_SECRET", "-v", "base64content==", "--file")
    assert r.exit_code == 0 and "Created personal secret 'FILE_SECRET'" in r.output and srv.writes[-1][2].get("isFile") is True
    for ok in ("API_KEY_2", "X"):
        assert run("create", "-n", ok, "-v", "value").exit_code == 0


@pytest.mark.parametrize("argv, typed", [((), "\n"), (("-n", "MY_SECRET"), "\n")])
def test_create_cancelled_at_a_prompt(hub, argv, typed):
    srv = hub()
    r = run("create", *argv, input=typed)
    assert r.exit_code == 0 and "Cancelled" in r.output and not srv.writes


@pytest.mark.parametrize("name", ["my_secret", "2FAST", "MY-SECRET", "lowercase_bad"])
def test_create_rejects_bad_names(hub, name):
    srv = hub()
    r = run("create", "-n", name, "-v", "value")
    assert r.exit_code != 0 and "Invalid secret name" in r.output and not srv.writes


def test_update(hub):
    srv = hub()
    r = run("update", SID, "-n", "RENAMED_SECRET")
    assert r.exit_code == 0 and "Updated secret" in r.output and srv.writes[-1][2] == {"name": "RENAMED_SECRET"}
    r = run("update", SID, "-v", "new-value")
    assert r.exit_code == 0 and "Updated secret" in r.output and srv.writes[-1][2] == {"value": "new-value"}
    n = len(srv.writes)
    r = run("update", SID, input="\n")
    assert r.exit_code == 0 and "No changes made" in r.output and len(srv.writes) == n
    r = run("update", SID, input="new-secret-value\n")
    assert r.exit_code == 0 and "Updated secret" in r.output and srv.writes[-1][2] == {"value": "new-secret-value"}
    assert "id" in json.loads(run("update", SID, "-n", "NEW_NAME", "-o", "json").output)


def test_delete_and_get(hub):
    srv = hub()
    r = run("delete", SID, "-y")
    assert r.exit_code == 0 and "Deleted secret" in r.output and srv.writes[-1][0] == "DELETE"
    r = run("delete", SID, input="n\n")
    assert r.exit_code == 0 and "Cancelled" in r.output and len(srv.writes) == 1
    r = run("delete", SID, input="y\n")
    assert r.exit_code == 0 and "Deleted secret" in r.output
    r = run("get", SID)
    assert r.exit_code == 0 and "Secret Details" in r.output and "MY_SECRET" in r.output
    doc = json.loads(run("get", SID, "-o", "json").output)
    assert doc["name"] == "MY_SECRET" and doc["id"] == SID


def test_help_texts():
    out = run("--help").output
    assert "Manage global secrets" in out and all(v in out for v in ("list", "create", "update", "delete", "get"))
    assert "--output" in run("list", "--help").output
    assert all(f in run("create", "--help").output for f in ("--name", "--value", "--description", "--file"))
    assert all(f in run("update", "--help").output for f in ("--name", "--value", "--description"))
    assert "--yes" in run("delete", "--help").output and "SECRET_ID" in run("get", "--help").output
