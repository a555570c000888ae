"""``prime switch [personal|<team slug>|<team id>]`` (reference: packages/prime/src/prime_cli/commands/switch.py:63-132)."""

from __future__ import annotations

from typing import Optional

import typer

from ..core import Config
from ._common import api, console, fail, handle_errors, make_app
from .teams import fetch_teams

app = make_app("Switch between personal and team accounts", invoke_without_command=True)
PERSONAL = "personal"


def find_team(teams: list[dict], target: str) -> dict | None:
    """Slug match wins over id match."""
    t = target.strip().lower()
    for key in ("slug", "teamId"):
        for team in teams:
            if str(team.get(key) or "").strip().lower() == t:
                return team
    return None


def _to_personal(c: Config) -> None:
    c.set_team(None)
    c.update_current_environment_file()
    console.print("[green]Switched to personal account.[/green]")


def _to_team(c: Config, team: dict) -> None:
    if not team.get("teamId"):
        raise fail("Selected team is missing a team ID.")
    c.set_team(team["teamId"], team_name=team.get("name", "Unknown"), team_role=team.get("role", "member"))
    c.update_current_environment_file()
    console.print(f"[green]Switched to team '{team.get('name', 'Unknown')}'.[/green]")


@app.callback(invoke_without_command=True)
@handle_errors
def switch(target: Optional[str] = typer.Argument(None, help=f"'{PERSONAL}', a team slug, or a team ID")) -> None:
    """Switch the active account (interactive when no target is given)."""
    c = Config()
    if c.team_id_from_env:
        raise fail("PRIME_TEAM_ID is set in your environment. Clear it before using [bold]prime switch[/bold].")
    if target is not None and target.strip().lower() == PERSONAL:
        return _to_personal(c)
    teams = fetch_teams(api())
    if target is not None:
        team = find_team(teams, target)
        if team is None:
            console.print(f"[red]Team '{target}' not found.[/red]")
            slugs = sorted(str(t["slug"]).strip() for t in teams if t.get("slug"))
            if slugs:
                console.print(f"[dim]Available teams: {', '.join(slugs)}[/dim]")
            raise typer.Exit(1)
        return _to_team(c, team)
    console.print("\n[bold]Switch account:[/bold]\n")
    console.print("  [cyan](1)[/cyan] Personal" + (" [green](current)[/green]" if c.team_id is None else ""))
    for i, t in enumerate(teams, 2):
        slug = str(t.get("slug") or "").strip()
        role = str(t.get("role", "member")).lower()
        detail = f"slug: {slug}, role: {role}" if slug else f"role: {role}"
        cur = " [green](current)[/green]" if t.get("teamId") == c.team_id else ""
        console.print(f"  [cyan]({i})[/cyan] {t.get('name', 'Unknown')} [dim]({detail})[/dim]{cur}")
    while True:
        n = typer.prompt("Select", type=int, default=1)
        if n == 1:
            return _to_personal(c)
        if 2 <= n <= len(teams) + 1:
            return _to_team(c, teams[n - 2])
        console.print(f"[red]Invalid selection. Enter 1-{len(teams) + 1}.[/red]")
