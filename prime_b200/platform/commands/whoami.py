"""``prime whoami`` — who the stored key authenticates as, which team context is active, what the token may do.
Also refreshes ``user_id`` in the config (same side effect and tables as the reference:
packages/prime/src/prime_cli/commands/whoami.py:15-88). ``--output json`` is ours: the same facts as one document.
"""

from __future__ import annotations

from typing import Any

from ..core import Config
from ..utils.display import build_table, output_data_as_json, validate_output_format
from ._common import OUTPUT_OPT, api, console, fail, handle_errors, make_app

app = make_app("Show the authenticated identity", invoke_without_command=True)

USER_FIELDS = (("User ID", "id", "Unknown"), ("Username", "slug", "[dim]Not set[/dim]"), ("Name", "name", "Unknown"), ("Email", "email", "Unknown"))


def account_sections(user: dict[str, Any], cfg: Config) -> list[list[tuple[str, str]]]:
    """Rows of the Account table, one inner list per ruled-off section: account type, team (if any), user."""
    sections: list[list[tuple[str, str]]] = [[("Type", "Team" if cfg.team_id else "Personal")]]
    if cfg.team_id:
        team = [("Team ID", cfg.team_id), ("Team Name", cfg.team_name or "[dim]Unknown[/dim]")]
        if cfg.team_role:
            team.append(("Role", cfg.team_role))
        sections.append(team)
    sections.append([(label, user.get(key) or missing) for label, key, missing in USER_FIELDS])
    return sections


def permission_rows(scope: dict[str, Any]) -> list[tuple[str, str, str]]:
    """``{"pods": {"read": true, "write": false}, "billing": null}`` → (name, ✓/✗, ✓/✗); a null scope shows dashes."""

    def mark(perms: dict[str, Any] | None, what: str) -> str:
        return "-" if perms is None else ("✓" if perms.get(what) else "✗")

    return [(name, mark(perms, "read"), mark(perms, "write")) for name, perms in scope.items()]


@app.callback(invoke_without_command=True)
@handle_errors
def whoami(output: str = OUTPUT_OPT) -> None:
    """Fetch identity, remember the user id, show account + token scopes."""
    user = api().get("/user/whoami").get("data")
    if not isinstance(user, dict):
        raise fail("Unexpected response from whoami endpoint")
    cfg = Config()
    if user.get("id"):
        cfg.set_user_id(user["id"])
        cfg.update_current_environment_file()
    scope = user.get("scope") or {}
    if validate_output_format(output, console) == "json":
        team = {"id": cfg.team_id, "name": cfg.team_name, "role": cfg.team_role} if cfg.team_id else None
        output_data_as_json({"user": {k: user.get(k) for _, k, _ in USER_FIELDS}, "team": team, "scope": scope}, console)
        return
    account = build_table("Account", [("Field", "cyan"), ("Value", "green")])
    for i, section in enumerate(account_sections(user, cfg)):
        if i:
            account.add_section()
        for row in section:
            account.add_row(*row)
    console.print(account)
    if scope:
        perms = build_table("Token Permissions", [("Scope", "cyan"), "Read", "Write"], permission_rows(scope))
        for col in perms.columns[1:]:
            col.justify = "center"
        console.print()
        console.print(perms)
