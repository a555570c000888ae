"""Behavioural conformance of ``prime env secret …`` with the reference CLI, scenario by scenario: same argv and stdin, same
exit status, same key phrases / JSON shape (scenarios from packages/prime/tests/test_env_secret.py:116-580; the harness and
the fake server are ours — routed on (method, path suffix), class-level so every command module sees it)."""

import json

import pytest
from typer.testing import CliRunner

from prime_b200.platform.core import client as core_client
from prime_b200.platform.main import app

runner = CliRunner()
ENV = "testuser/test-env"
SECRETS = [
    {"id": "esecret-id-001", "name": "DB_PASSWORD", "source": "environment", "description": "Database password",
     "createdAt": "2026-01-15T10:00:00Z", "updatedAt": "2026-01-15T10:00:00Z"},
    {"id": "esecret-id-002", "name": "OPENAI_KEY", "source": "global-linked", "description": None,
     "createdAt": "2026-01-10T08:00:00Z", "updatedAt": "2026-01-10T08:00:00Z"},
]  # fmt: skip


class Server:
    """What the hub would answer; remembers the writes."""

    def __init__(self, secrets):
        self.secrets, self.writes = list(secrets), []

    def get(self, endpoint, params=None, **kw):
        if "/@latest" in endpoint:
            return {"data": {"id": "env-uuid-12345", "name": "test-env", "owner": {"name": "testuser", "type": "user"}}}
        if endpoint.endswith("/secrets"):
            return {"data": self.secrets}
        if "/secrets/" in endpoint:
            return {"data": next((s for s in self.secrets if endpoint.endswith(s["id"])), self.secrets[0])}
        return {"data": {}}

    def post(self, endpoint, json=None, **kw):
        self.writes.append(("POST", endpoint, json))
        if "/secrets/link/" in endpoint:
            return {"data": {"id": "link-id-12345", "secretId": endpoint.rsplit("/", 1)[1], "secretName": "LINKED_SECRET", "environmentId": "env-uuid-12345"}}
        return {"data": {"id": "new-esecret-id-001", "name": (json or {}).get("name"), "value": "[encrypted]",
                         "description": (json or {}).get("description"), "source": "environment"}}  # fmt: skip

    def patch(self, endpoint, json=None, **kw):
        self.writes.append(("PATCH", endpoint, json))
        return {"data": {"id": "esecret-id-001", "name": (json or {}).get("name", "DB_PASSWORD"), "description": (json or {}).get("description"), "source": "environment"}}

    def delete(self, endpoint, **kw):
        self.writes.append(("DELETE", endpoint, None))
        return {}


@pytest.fixture
def hub(monkeypatch):
    def install(secrets=SECRETS):
        srv = Server(secrets)
        monkeypatch.setenv("PRIME_API_KEY", "test-key")
        for verb in ("get", "post", "patch", "delete"):
            monkeypatch.setattr(core_client.APIClient, verb, lambda self, *a, _v=verb, **kw: getattr(srv, _v)(*a, **kw))
        return srv

    return install


def run(*argv, input=None):
    return runner.invoke(app, ["env", "secret", *argv], input=input)


def test_list(hub):
    hub()
    r = run("list", ENV)
    assert r.exit_code == 0 and f"Secrets for {ENV}" in r.output and "DB_PASSWORD" in r.output and "OPENAI_KEY" in r.output
    assert "environment" in r.output and "global-linked" in r.output  # the source column
    doc = json.loads(run("list", ENV, "-o", "json").output)
    assert len(doc["secrets"]) == 2 and doc["secrets"][0]["name"] == "DB_PASSWORD"
    r = run("list", ENV, "-o", "xml")
    assert r.exit_code != 0 and "Invalid output format" in r.output


def test_list_empty(hub):
    hub(secrets=[])
    r = run("list", ENV)
    assert r.exit_code == 0 and "No secrets found" in r.output


def test_create(hub):
    srv = hub()
    r = run("create", ENV, "-n", "NEW_SECRET", "-v", "secret-value", "-d", "A test secret")
    assert r.exit_code == 0 and "Created secret 'NEW_SECRET'" in r.output and ENV in r.output and "ID:" in r.output
    assert srv.writes[-1][2] == {"name": "NEW_SECRET", "value": "secret-value", "description": "A test secret"}
    doc = json.loads(run("create", ENV, "-n", "NEW_SECRET", "-v", "value", "-o", "json").output)
    assert doc["name"] == "NEW_SECRET" and "id" in doc
    r = run("create", ENV, input="MY_NEW_SECRET\nsecret-value\n")
    assert r.exit_code == 0 and "Created secret" in r.output


@pytest.mark.parametrize("argv, typed", [((), "\n"), (("-n", "NEW_SECRET"), "\n")])
def test_create_cancelled_at_a_prompt(hub, argv, typed):
    srv = hub()
    r = run("create", ENV, *argv, input=typed)
    assert r.exit_code == 0 and "Cancelled" in r.output and not srv.writes


@pytest.mark.parametrize("name", ["my_secret", "1BAD_NAME", "MY-SECRET"])
def test_create_rejects_bad_names(hub, name):
    srv = hub()
    r = run("create", ENV, "-n", name, "-v", "value")
    assert r.exit_code != 0 and "Invalid secret name" in r.output and not srv.writes


@pytest.mark.parametrize("change, body", [(("-n", "RENAMED"), {"name": "RENAMED"}), (("-v", "new-value"), {"value": "new-value"}),
                                          (("-d", "Updated description"), {"description": "Updated description"})])  # fmt: skip
def test_update_by_id(hub, change, body):
    srv = hub()
    r = run("update", ENV, "--id", "esecret-id-001", *change)
    assert r.exit_code == 0 and "Updated secret" in r.output
    assert srv.writes[-1][0] == "PATCH" and srv.writes[-1][1].endswith("/secrets/esecret-id-001") and srv.writes[-1][2] == body


def test_update_other_paths(hub):
    srv = hub()
    assert "id" in json.loads(run("update", ENV, "--id", "esecret-id-001", "-n", "RENAMED", "-o", "json").output)
    r = run("update", ENV, "--id", "esecret-id-001", input="\n")  # nothing given, empty value at the prompt
    assert r.exit_code == 0 and "No changes made" in r.output
    n = len(srv.writes)
    r = run("update", ENV, input="1\nnew-secret-value\n")  # pick from the numbered list, then type the value
    assert r.exit_code == 0 and "Updated secret" in r.output and srv.writes[n][2] == {"value": "new-secret-value"}
    r = run("update", ENV, input="\n")
    assert r.exit_code == 0 and "Cancelled" in r.output and len(srv.writes) == n + 1


def test_delete(hub):
    srv = hub()
    r = run("delete", ENV, "--id", "esecret-id-001", "-y")
    assert r.exit_code == 0 and "Deleted secret" in r.output and ENV in r.output and srv.writes[-1][0] == "DELETE"
    n = len(srv.writes)
    r = run("delete", ENV, "--id", "esecret-id-001", input="n\n")
    assert r.exit_code == 0 and "Cancelled" in r.output and len(srv.writes) == n
    r = run("delete", ENV, input="1\ny\n")
    assert r.exit_code == 0 and "Deleted secret" in r.output and len(srv.writes) == n + 1
    r = run("delete", ENV, input="\n")
    assert r.exit_code == 0 and "Cancelled" in r.output


def test_delete_with_nothing_to_delete(hub):
    hub(secrets=[])
    r = run("delete", ENV)
    assert r.exit_code == 0 and "No secrets to delete" in r.output


def test_link_and_unlink(hub):
    srv = hub()
    r = run("link", "global-secret-id-123", ENV)
    assert r.exit_code == 0 and "Linked global secret" in r.output and ENV in r.output
    assert srv.writes[-1][1].endswith("/secrets/link/global-secret-id-123")
    doc = json.loads(run("link", "global-secret-id-123", ENV, "-o", "json").output)
    assert "secretId" in doc or "id" in doc
    r = run("unlink", "global-secret-id-123", ENV, "-y")
    assert r.exit_code == 0 and "Unlinked global secret" in r.output and ENV in r.output and srv.writes[-1][0] == "DELETE"
    n = len(srv.writes)
    r = run("unlink", "global-secret-id-123", ENV, input="n\n")
    assert r.exit_code == 0 and "Cancelled" in r.output and len(srv.writes) == n
    r = run("unlink", "global-secret-id-123", ENV, input="y\n")
    assert r.exit_code == 0 and "Unlinked global secret" in r.output


def test_help_texts():
    out = run("--help").output
    assert all(v in out for v in ("list", "create", "update", "delete", "link", "unlink"))
    assert "--output" in run("list", "--help").output
    assert all(f in run("create", "--help").output for f in ("--name", "--value", "--description"))
    assert "GLOBAL_SECRET_ID" in run("link", "--help").output and "GLOBAL_SECRET_ID" in run("unlink", "--help").output
