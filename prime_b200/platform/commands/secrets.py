"""``prime secret {list,create,update,delete,get}`` — global (personal or team) secrets
(reference: packages/prime/src/prime_cli/commands/secrets.py:43-343; endpoints /secrets/, /secrets/{id})."""

from __future__ import annotations

from typing import Any, Optional

import typer

from ..core import Config
from ..utils.display import output_data_as_json
from ..utils.json_help import json_output_help, list_json_help
from ..utils.prompt import any_provided, confirm_or_skip, prompt_for_value, require_selection, validate_env_var_name
from ..utils.time_utils import format_time_ago
from ._common import OUTPUT_OPT, api, console, emit, handle_errors, make_app

app = make_app("Manage global secrets")
_FIELDS = {"id": "str", "name": "str", "description": "str|null", "isFile": "bool", "createdAt": "str", "updatedAt": "str"}
DETAIL_HELP = json_output_help(_FIELDS)


def team_params(config: Config) -> dict[str, str] | None:
    return {"teamId": config.team_id} if config.team_id else None


def scope_of(config: Config) -> str:
    return "team" if config.team_id else "personal"


def fetch_secrets(client, config: Config) -> list[dict[str, Any]]:
    return client.get("/secrets/", params=team_params(config)).get("data", [])


def _cancel() -> "typer.Exit":
    console.print("\n[dim]Cancelled.[/dim]")
    return typer.Exit()


@app.command("list", epilog=list_json_help("secrets", _FIELDS))
@handle_errors
def secret_list(output: str = OUTPUT_OPT) -> None:
    """List secrets of the active account (values are never returned)."""
    cfg = Config(writable=False)
    rows = fetch_secrets(api(), cfg)
    if not rows and output == "table":
        console.print(f"[yellow]No {scope_of(cfg)} secrets found.[/yellow]\n[dim]Create one with: prime secret create[/dim]")
        return
    emit(output, {"secrets": rows, "total_count": len(rows)}, f"{scope_of(cfg).capitalize()} Secrets",
         [("ID", "cyan"), ("Name", "green"), "Description", "File", ("Updated", "magenta")],
         [[s.get("id"), s.get("name"), s.get("description") or "", "yes" if s.get("isFile") else "", format_time_ago(s.get("updatedAt"))] for s in rows])  # fmt: skip


@app.command("create", epilog=DETAIL_HELP)
@handle_errors
def secret_create(
    name: Optional[str] = typer.Option(None, "--name", "-n", help="Secret name (becomes the environment variable name)"),
    value: Optional[str] = typer.Option(None, "--value", "-v", help="Secret value (prompted, hidden, when omitted)"),
    description: Optional[str] = typer.Option(None, "--description", "-d"),
    is_file: bool = typer.Option(False, "--file", "-f", help="Value is base64 file content"),
    output: str = OUTPUT_OPT,
) -> None:
    """Create a secret."""
    name = name or prompt_for_value("Secret name")
    if not name:
        raise _cancel()
    if not validate_env_var_name(name, "secret"):
        raise typer.Exit(1)
    value = value or prompt_for_value("Secret value", hide_input=True)
    if not value:
        raise _cancel()
    cfg = Config(writable=False)
    body: dict[str, Any] = {"name": name, "value": value}
    if description:
        body["description"] = description
    if is_file:
        body["isFile"] = True
    if cfg.team_id:
        body["teamId"] = cfg.team_id
    secret = api().post("/secrets/", json=body).get("data", {})
    if output == "json":
        return output_data_as_json(secret, console)
    console.print(f"[green]✓ Created {scope_of(cfg)} secret '{name}'[/green]\n[dim]ID: {secret.get('id')}[/dim]")


@app.command("update", epilog=DETAIL_HELP)
@handle_errors
def secret_update(
    secret_id: Optional[str] = typer.Argument(None, help="Secret ID (interactive pick when omitted)"),
    name: Optional[str] = typer.Option(None, "--name", "-n"),
    value: Optional[str] = typer.Option(None, "--value", "-v"),
    description: Optional[str] = typer.Option(None, "--description", "-d"),
    output: str = OUTPUT_OPT,
) -> None:
    """Update name, value or description of a secret."""
    client, cfg = api(), Config(writable=False)
    if not secret_id:
        secret_id = require_selection(fetch_secrets(client, cfg), "update", f"No {scope_of(cfg)} secrets to update.", "secret").get("id")
    if not any_provided(name, value, description):
        console.print("\n[bold]What would you like to update?[/bold]")
        value = prompt_for_value("New value", required=False, hide_input=True) or None
        if not value:
            console.print("\n[dim]No changes made.[/dim]")
            raise typer.Exit()
    if name is not None and not validate_env_var_name(name, "secret"):
        raise typer.Exit(1)
    body = {k: v for k, v in (("name", name), ("value", value), ("description", description)) if v is not None}
    secret = client.patch(f"/secrets/{secret_id}", json=body, params=team_params(cfg)).get("data", {})
    if output == "json":
        return output_data_as_json(secret, console)
    console.print(f"[green]✓ Updated secret '{secret.get('name', secret_id)}'[/green]")


@app.command("delete")
@handle_errors
def secret_delete(secret_id: Optional[str] = typer.Argument(None, help="Secret ID (interactive pick when omitted)"),
                  yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation")) -> None:  # fmt: skip
    """Delete a secret."""
    client, cfg = api(), Config(writable=False)
    if not secret_id:
        chosen = require_selection(fetch_secrets(client, cfg), "delete", f"No {scope_of(cfg)} secrets to delete.", "secret")
        secret_id, label = chosen.get("id"), chosen.get("name")
    else:
        label = client.get(f"/secrets/{secret_id}", params=team_params(cfg)).get("data", {}).get("name", secret_id)
    if not confirm_or_skip(f"Delete secret '{label}'?", yes):
        raise _cancel()
    client.delete(f"/secrets/{secret_id}", params=team_params(cfg))
    console.print(f"[green]✓ Deleted secret '{label}'[/green]")


@app.command("get", epilog=DETAIL_HELP)
@handle_errors
def secret_get(secret_id: str = typer.Argument(..., help="Secret ID"), output: str = OUTPUT_OPT) -> None:
    """Show a secret's metadata."""
    s = api().get(f"/secrets/{secret_id}", params=team_params(Config(writable=False))).get("data", {})
    emit(output, s, "Secret Details", [("Field", "cyan"), ("Value", "green")],
         [["ID", s.get("id")], ["Name", s.get("name")], ["Description", s.get("description") or ""], ["File", bool(s.get("isFile"))],
          ["Created", s.get("createdAt")], ["Updated", s.get("updatedAt")]])  # fmt: skip
