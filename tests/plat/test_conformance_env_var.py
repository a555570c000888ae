"""Behavioural conformance of ``prime env var …`` with the reference CLI (scenarios: packages/prime/tests/test_env_var.py:101-440;
harness and fake hub are ours — see test_conformance_env_secret.py)."""

import json

import pytest
from typer.testing import CliRunner

from prime_b200.platform.core import client as core_client
from prime_b200.platform.main import app

runner = CliRunner()
ENV = "testuser/test-env"
VID = "var-id-1234567890"
VARS = [
    {"id": VID, "name": "DEBUG", "value": "true", "description": "Enable debug mode", "createdAt": "2026-01-15T10:00:00Z", "updatedAt": "2026-01-15T10:00:00Z"},
    {"id": "var-id-0987654321", "name": "LOG_LEVEL", "value": "info", "description": None, "createdAt": "2026-01-10T08:00:00Z", "updatedAt": "2026-01-10T08:00:00Z"},
]  # fmt: skip


class Server:
    def __init__(self, variables):
        self.variables, self.writes = list(variables), []

    def get(self, endpoint, params=None, **kw):
        if "/@latest" in endpoint:
            return {"data": {"id": "env-uuid-12345", "name": "test-env", "owner": {"name": "testuser", "type": "user"}}}
        if endpoint.endswith("/variables"):
            return {"data": self.variables}
        if "/variables/" in endpoint:
            tail = endpoint.rsplit("/", 1)[1]
            return {"data": next((v for v in self.variables if v["id"].startswith(tail)), self.variables[0])}
        return {"data": {}}

    def post(self, endpoint, json=None, **kw):
        self.writes.append(("POST", endpoint, json))
        return {"data": {"id": "new-var-id-001", **(json or {})}}

    def patch(self, endpoint, json=None, **kw):
        self.writes.append(("PATCH", endpoint, json))
        return {"data": {**self.variables[0], **(json or {})}}

    def delete(self, endpoint, **kw):
        self.writes.append(("DELETE", endpoint, None))
        return {}


@pytest.fixture
def hub(monkeypatch):
    def install(variables=VARS):
        srv = Server(variables)
        monkeypatch.setenv("PRIME_API_KEY", "test-key")
        for verb in ("get", "post", "patch", "delete"):
            monkeypatch.setattr(core_client.APIClient, verb, lambda self, *a, _v=verb, **kw: getattr(srv, _v)(*a, **kw))
        return srv

    return install


def run(*argv, input=None):
    return runner.invoke(app, ["env", "var", *argv], input=input)


def test_list(hub):
    hub()
    r = run("list", ENV)
    assert r.exit_code == 0 and f"Variables for {ENV}" in r.output and all(x in r.output for x in ("DEBUG", "LOG_LEVEL", "true"))
    doc = json.loads(run("list", ENV, "-o", "json").output)
    assert len(doc["variables"]) == 2 and doc["variables"][0]["name"] == "DEBUG"
    r = run("list", ENV, "-o", "xml")
    assert r.exit_code != 0 and "Invalid output format" in r.output


def test_list_empty(hub):
    hub(variables=[])
    r = run("list", ENV)
    assert r.exit_code == 0 and "No variables found" in r.output


def test_create(hub):
    srv = hub()
    r = run("create", ENV, "-n", "NEW_VAR", "-v", "value")
    assert r.exit_code == 0 and "Created variable 'NEW_VAR'" in r.output and "ID:" in r.output
    r = run("create", ENV, "-n", "NEW_VAR", "-v", "value", "-d", "A test variable")
    assert r.exit_code == 0 and srv.writes[-1][2] == {"name": "NEW_VAR", "value": "value", "description": "A test variable"}
    doc = json.loads(run("create", ENV, "-n", "NEW_VAR", "-v", "value", "-o", "json").output)
    assert doc["name"] == "NEW_VAR" and "id" in doc
    r = run("create", ENV, input="NEW_VAR\nmy-value\n")
    assert r.exit_code == 0 and "Created variable 'NEW_VAR'" in r.output
    assert run("create", ENV, "-n", "VAR_2", "-v", "value").exit_code == 0  # digits are fine after the first character
    r = run("create", ENV, "-n", "NEW_VAR", "-v", "value", "-o", "xml")
    assert r.exit_code != 0 and "Invalid output format" in r.output


@pytest.mark.parametrize("argv, typed", [((), "\n"), (("-n", "NEW_VAR"), "\n")])
def test_create_cancelled_at_a_prompt(hub, argv, typed):
    srv = hub()
    r = run("create", ENV, *argv, input=typed)
    assert r.exit_code == 0 and "Cancelled" in r.output and not srv.writes


@pytest.mark.parametrize("name", ["lowercase", "3RD_VAR", "MY-VAR"])
def test_create_rejects_bad_names(hub, name):
    srv = hub()
    r = run("create", ENV, "-n", name, "-v", "value")
    assert r.exit_code != 0 and "Invalid variable name" in r.output and not srv.writes


@pytest.mark.parametrize("change, body", [
    (("-n", "RENAMED_VAR"), {"name": "RENAMED_VAR"}), (("-v", "false"), {"value": "false"}), (("-d", "New description"), {"description": "New description"}),
    (("-n", "RENAMED", "-v", "new-val", "-d", "all three"), {"name": "RENAMED", "value": "new-val", "description": "all three"}),
])  # fmt: skip
def test_update(hub, change, body):
    srv = hub()
    r = run("update", VID, ENV, *change)
    assert r.exit_code == 0 and "Updated variable" in r.output
    assert srv.writes[-1][0] == "PATCH" and srv.writes[-1][1].endswith(f"/variables/{VID}") and srv.writes[-1][2] == body


def test_update_needs_a_change_and_speaks_json(hub):
    srv = hub()
    r = run("update", VID, ENV)
    assert r.exit_code == 1 and "At least one of --name, --value, or --description is required" in r.output and not srv.writes
    assert "id" in json.loads(run("update", VID, ENV, "-n", "RENAMED_VAR", "-o", "json").output)


def test_delete(hub):
    srv = hub()
    r = run("delete", VID, ENV, "-y")
    assert r.exit_code == 0 and "Variable deleted" in r.output and srv.writes[-1][0] == "DELETE"
    r = run("delete", VID, ENV, input="n\n")
    assert r.exit_code == 0 and "Cancelled" in r.output and len(srv.writes) == 1
    r = run("delete", VID, ENV, input="y\n")
    assert r.exit_code == 0 and "Variable deleted" in r.output and len(srv.writes) == 2
