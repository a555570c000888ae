"""Conformance of ``prime config`` with the reference: the team-id shape check (CUID v1: 'c' + 24 lower-case alphanumerics, or
empty for the personal account) and ``set-team-id`` resolving the team's name through a PAGINATED team listing
(scenarios: packages/prime/tests/test_config.py:7-60, test_config_set_team.py:22-80; harness is ours)."""

import pytest
from typer.testing import CliRunner

from prime_b200.platform.commands.config import validate_team_id
from prime_b200.platform.core import Config
from prime_b200.platform.core import client as core_client
from prime_b200.platform.main import app

runner = CliRunner()


@pytest.mark.parametrize("team_id, ok", [
    ("cmf0ohr9s0026ilerf3w68s6n", True), ("", True),
    ("CMF0OHR9S0026ILERF3W68S6N", False), ("CmF0OhR9s0026IlErF3w68S6n", False),  # CUID v1 is lower case only
    ("amf0ohr9s0026ilerf3w68s6n", False),  # must start with 'c'
    ("cmf0ohr9s0026ilerf3w68s6", False), ("cmf0ohr9s0026ilerf3w68s6nn", False),  # exactly 25 characters
    ("cmf0ohr9s0026ilerf3w68s6!", False), ("cmf0ohr9s0026 lerf3w68s6n", False), ("team", False), ("cmf0ohr9s0026_lerf3w68s6n", False),
])  # fmt: skip
def test_team_id_shape(team_id, ok):
    assert validate_team_id(team_id) is ok


def test_set_team_id_wants_its_argument():
    r = runner.invoke(app, ["config", "set-team-id"])
    assert r.exit_code != 0 and "Missing argument 'TEAM_ID'" in r.output


def test_set_team_id_finds_the_name_on_the_second_page(monkeypatch):
    monkeypatch.setenv("PRIME_API_KEY", "test-key")
    target = {"teamId": "cmf0ohr9s0026ilerf3w68s6z", "name": "Page Two Team", "slug": "page-two", "role": "admin", "createdAt": "2026-01-15T10:00:00Z"}
    first_page = [{"teamId": f"cmf0ohr9s0026ilerf3w68{i:02d}", "name": f"Team {i}", "slug": f"team-{i}", "role": "member", "createdAt": "2026-01-15T10:00:00Z"}
                  for i in range(100)]  # fmt: skip
    asked = []

    def get(self, endpoint, params=None, **kw):
        if endpoint == "/user/teams":
            offset, limit = (params or {}).get("offset", 0), (params or {}).get("limit", 100)
            asked.append(offset)
            if offset == 0:
                return {"data": first_page[:limit], "total_count": 101}
            if offset == 100:
                return {"data": [target], "total_count": 101}
        return {"data": []}

    monkeypatch.setattr(core_client.APIClient, "get", get)
    r = runner.invoke(app, ["config", "set-team-id", target["teamId"]])
    assert r.exit_code == 0, r.output
    assert f"Team '{target['name']}' ({target['teamId']}) configured successfully!" in r.output
    assert asked == [0, 100]
    cfg = Config(writable=False)
    assert cfg.team_id == target["teamId"] and cfg.team_name == "Page Two Team" and cfg.team_role == "admin"
