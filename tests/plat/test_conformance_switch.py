"""Behavioural conformance of ``prime switch`` with the reference CLI: personal / slug / id / interactive, unknown teams, teams
without a slug, and the PRIME_TEAM_ID guard (scenarios: packages/prime/tests/test_switch.py:71-213; harness is ours)."""

import io
import json

import pytest
from typer.testing import CliRunner

from prime_b200.platform.commands import switch as switch_mod
from prime_b200.platform.core import Config
from prime_b200.platform.core import client as core_client
from prime_b200.platform.main import app
from prime_b200.platform.utils.plain import get_console

runner = CliRunner()
T1, T2 = "cmf0ohr9s0026ilerf3w68s6n", "cmf0ohr9s0026ilerf3w68s6m"
TEAMS = [{"teamId": T1, "name": "Prime Team", "slug": "prime", "role": "admin", "createdAt": "2026-01-15T10:00:00Z"},
         {"teamId": T2, "name": "Research Team", "slug": "research", "role": "member", "createdAt": "2026-01-15T10:00:00Z"}]  # fmt: skip
NO_SLUG = [{"teamId": T1, "name": "New team", "slug": None, "role": "ADMIN", "createdAt": "2026-01-15T10:00:00Z"}]


@pytest.fixture
def hub(monkeypatch):
    def install(teams=TEAMS):
        monkeypatch.setenv("PRIME_API_KEY", "test-key")

        def get(self, endpoint, params=None, **kw):
            if endpoint == "/user/teams":
                return {"data": teams}
            if endpoint == "/user/whoami":
                return {"data": {"id": "user-123"}}
            return {"data": []}

        monkeypatch.setattr(core_client.APIClient, "get", get)

    return install


def test_console_factory_takes_plain_rendering_arguments():
    assert get_console(file=io.StringIO(), markup=False, highlight=False, no_color=True, emoji=False) is not None


def test_personal(hub):
    hub()
    for argv in (["switch", "personal"], ["switch", "--plain", "personal"]):
        r = runner.invoke(app, argv)
        assert r.exit_code == 0 and "Switched to personal account." in r.output, r.output
    assert Config(writable=False).team_id is None


def test_personal_needs_neither_key_nor_team_listing(monkeypatch):
    monkeypatch.setattr(switch_mod, "fetch_teams", lambda *_a, **_k: pytest.fail("the team list is not needed to go personal"))
    r = runner.invoke(app, ["switch", "personal"])
    assert r.exit_code == 0 and "Switched to personal account." in r.output, r.output


@pytest.mark.parametrize("ref", ["prime", T1])
def test_team_by_slug_or_id(hub, ref):
    hub()
    r = runner.invoke(app, ["switch", ref])
    assert r.exit_code == 0 and "Switched to team 'Prime Team'." in r.output, r.output
    cfg = Config(writable=False)
    assert cfg.team_id == T1 and cfg.team_name == "Prime Team"


def test_unknown_team_lists_what_exists(hub):
    hub()
    r = runner.invoke(app, ["switch", "unknown"])
    assert r.exit_code == 1 and "Team 'unknown' not found." in r.output and "Available teams: prime, research" in r.output


def test_a_missing_slug_never_matches_the_word_none(hub):
    hub(NO_SLUG)
    r = runner.invoke(app, ["switch", "none"])
    assert r.exit_code == 1 and "Team 'none' not found." in r.output


def test_interactive(hub):
    hub()
    r = runner.invoke(app, ["switch"], input="2\n")
    assert r.exit_code == 0 and "Switch account:" in r.output and "Prime Team (slug: prime, role: admin)" in r.output
    assert "Switched to team 'Prime Team'." in r.output


def test_interactive_marks_only_the_current_account_and_hides_missing_slugs(hub, isolated_home):
    hub(NO_SLUG)
    (isolated_home / ".prime" / "environments").mkdir(parents=True)
    (isolated_home / ".prime" / "config.json").write_text(json.dumps({
        "api_key": "", "team_id": T1, "team_name": "New team", "team_role": "ADMIN", "user_id": None, "base_url": "https://api.primeintellect.ai",
        "frontend_url": "https://app.primeintellect.ai", "inference_url": "https://api.pinference.ai/api/v1", "ssh_key_path": "~/.ssh/id_rsa",
        "current_environment": "production"}))  # fmt: skip
    r = runner.invoke(app, ["switch"], input="1\n")
    assert r.exit_code == 0, r.output
    assert "Personal (current)" not in r.output and "New team (role: admin) (current)" in r.output and "slug:" not in r.output


def test_refuses_while_the_environment_pins_a_team(hub, monkeypatch):
    hub()
    monkeypatch.setenv("PRIME_TEAM_ID", T1)
    r = runner.invoke(app, ["switch", "personal"])
    assert r.exit_code == 1 and "PRIME_TEAM_ID is set in your environment" in r.output
