"""``prime whoami`` (reference: packages/prime/src/prime_cli/commands/whoami.py:15-88)."""

from __future__ import annotations

from rich.table import Table

from ..core import Config
from ._common import api, console, fail, handle_errors, make_app

app = make_app("Show the authenticated identity", invoke_without_command=True)


@app.callback(invoke_without_command=True)
@handle_errors
def whoami() -> None:
    """Fetch identity, remember the user id, show account + token scopes."""
    data = api().get("/user/whoami").get("data")
    if not isinstance(data, dict):
        raise fail("Unexpected response from whoami endpoint")
    c = Config()
    if data.get("id"):
        c.set_user_id(data["id"])
        c.update_current_environment_file()
    t = Table(title="Account")
    t.add_column("Field", style="cyan")
    t.add_column("Value", style="green")
    if c.team_id:
        t.add_row("Type", "Team")
        t.add_section()
        t.add_row("Team ID", c.team_id)
        t.add_row("Team Name", c.team_name or "[dim]Unknown[/dim]")
        if c.team_role:
            t.add_row("Role", c.team_role)
    else:
        t.add_row("Type", "Personal")
    t.add_section()
    for label, key, missing in (("User ID", "id", "Unknown"), ("Username", "slug", "[dim]Not set[/dim]"), ("Name", "name", "Unknown"),
                                ("Email", "email", "Unknown")):  # fmt: skip
        t.add_row(label, data.get(key) or missing)
    console.print(t)
    scope = data.get("scope") or {}
    if scope:
        p = Table(title="Token Permissions")
        p.add_column("Scope", style="cyan")
        p.add_column("Read", justify="center")
        p.add_column("Write", justify="center")
        for name, perms in scope.items():
            if perms is None:
                p.add_row(name, "-", "-")
            else:
                p.add_row(name, "✓" if perms.get("read") else "✗", "✓" if perms.get("write") else "✗")
        console.print()
        console.print(p)
