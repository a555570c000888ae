"""``prime teams {list,members}`` (reference: packages/prime/src/prime_cli/commands/teams.py:24-165)."""

from __future__ import annotations

from typing import Optional

import typer

from ..core import Config
from ..utils.json_help import list_json_help
from ._common import OUTPUT_OPT, api, emit, fail, handle_errors, make_app, paginate_hint

app = make_app("Teams")


def fetch_teams(client, page: int = 100) -> list[dict]:
    """All teams of the user, following offset pagination."""
    teams: list[dict] = []
    while True:
        resp = client.get("/user/teams", params={"offset": len(teams), "limit": page})
        batch = resp.get("data", []) if isinstance(resp, dict) else []
        teams += batch
        if not batch or len(teams) >= resp.get("total_count", len(teams)):
            return teams


def fetch_team_members(client, team_id: str) -> list[dict]:
    resp = client.get(f"/teams/{team_id}/members")
    return resp.get("data", []) if isinstance(resp, dict) else []


@app.command("list", epilog=list_json_help("teams", {"teamId": "str", "name": "str", "slug": "str", "role": "str", "createdAt": "str"},
                                           {"total_count": "int", "offset": "int", "limit": "int"}))  # fmt: skip
@handle_errors
def list_teams(limit: int = typer.Option(100, help="Maximum number of teams"), offset: int = typer.Option(0, help="Teams to skip"),
               output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """List your teams."""
    resp = api().get("/user/teams", params={"offset": offset, "limit": limit})
    teams = resp.get("data", [])
    total = resp.get("total_count", len(teams))
    emit(output, {"teams": teams, "total_count": total, "offset": offset, "limit": limit}, f"Teams (Total: {total})",
         [("ID", "cyan"), ("Name", "blue"), ("Slug", "green"), ("Role", "yellow"), ("Created", "magenta")],
         [[t.get("teamId"), t.get("name"), t.get("slug"), t.get("role"), t.get("createdAt")] for t in teams],
         paginate_hint(total, offset, limit, "teams"))  # fmt: skip


@app.command("members")
@handle_errors
def list_members(team_id: Optional[str] = typer.Option(None, "--team-id", help="Defaults to the configured team"),
                 output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """List members of a team."""
    tid = team_id or Config(writable=False).team_id
    if not tid:
        raise fail("No team selected. Use --team-id or 'prime switch <team>'.")
    members = fetch_team_members(api(), tid)
    emit(output, {"members": members, "total_count": len(members)}, f"Team Members (Total: {len(members)})",
         [("User ID", "cyan"), ("Name", "blue"), ("Email", "green"), ("Role", "yellow"), ("Joined", "magenta")],
         [[m.get("userId", ""), m.get("userName") or "N/A", m.get("userEmail") or "N/A", m.get("role", ""), m.get("joinedAt", "")] for m in members])  # fmt: skip
