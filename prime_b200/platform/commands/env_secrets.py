"""``prime env secret`` / ``prime env var`` — per-environment secrets (write-only values, optional links to global secrets)
and plain variables. One implementation parameterised by kind
(reference: packages/prime/src/prime_cli/commands/env.py:3031-3659)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional

import typer

from ..core import APIClient
from ..utils.display import output_data_as_json, validate_output_format
from ..utils.prompt import any_provided, confirm_or_skip, prompt_for_value, require_selection, validate_env_var_name
from ..utils.time_utils import format_time_ago
from ._common import OUTPUT_OPT, api, console, emit, fail, handle_errors, make_app
from .env import app as env_app
from .env import get_environment_id, resolve_environment

secret_app = make_app("Manage environment secrets")
var_app = make_app("Manage environment variables")
env_app.add_typer(secret_app, name="secret", rich_help_panel="Manage")
env_app.add_typer(var_app, name="var", rich_help_panel="Manage")

ENV_ARG = typer.Argument(None, help="Environment slug (owner/name). Auto-detected from the current directory if omitted.")


@dataclass(frozen=True)
class Kind:
    noun: str  # "secret" | "variable"
    path: str  # URL segment
    json_key: str
    hidden: bool  # value never shown / prompted without echo


SECRET = Kind("secret", "secrets", "secrets", True)
VARIABLE = Kind("variable", "variables", "variables", False)


def fetch_items(client: APIClient, env_id: str, kind: Kind) -> list[dict[str, Any]]:
    return client.get(f"/environmentshub/{env_id}/{kind.path}").get("data", [])


def _cancelled() -> "typer.Exit":
    console.print("\n[dim]Cancelled.[/dim]")
    return typer.Exit()


def _list(kind: Kind, environment: Optional[str], output: str) -> None:
    validate_output_format(output, console)
    owner, name = resolve_environment(environment)
    client = api()
    items = fetch_items(client, get_environment_id(client, owner, name), kind)
    if not items and output != "json":
        console.print(f"[yellow]No {kind.noun}s found for this environment.[/yellow]")
        return
    third = ("Source", "blue") if kind.hidden else ("Value", "green")
    rows = []
    for it in items:
        mid = it.get("source", "") if kind.hidden else (it.get("value", "") if len(it.get("value", "")) <= 30 else it["value"][:27] + "...")
        rows.append([it.get("id", ""), it.get("name", ""), mid, it.get("description") or "", format_time_ago(it["createdAt"]) if it.get("createdAt") else ""])
    emit(output, {kind.json_key: items}, f"{kind.noun.capitalize()}s for {owner}/{name}",
         [("ID", "dim"), ("Name", "cyan"), third, ("Description", "dim"), ("Created", "dim")], rows)  # fmt: skip


def _create(kind: Kind, environment: Optional[str], name: Optional[str], value: Optional[str], description: Optional[str], output: str) -> None:
    validate_output_format(output, console)
    owner, env_name = resolve_environment(environment)
    name = name or prompt_for_value(f"{kind.noun.capitalize()} name")
    if not name:
        raise _cancelled()
    if not validate_env_var_name(name, kind.noun):
        raise typer.Exit(1)
    value = value or prompt_for_value(f"{kind.noun.capitalize()} value", hide_input=kind.hidden)
    if not value:
        raise _cancelled()
    payload: dict[str, Any] = {"name": name, "value": value}
    if description:
        payload["description"] = description
    with console.status(f"[bold blue]Creating {kind.noun}...", spinner="dots"):
        client = api()
        env_id = get_environment_id(client, owner, env_name)
        item = client.post(f"/environmentshub/{env_id}/{kind.path}", json=payload).get("data", {})
    if output == "json":
        output_data_as_json(item, console)
        return
    console.print(f"[green]✓ Created {kind.noun} '{name}' for {owner}/{env_name}[/green]")
    console.print(f"[dim]ID: {item.get('id')}[/dim]")


def _patch(kind: Kind, client: APIClient, env_id: str, item_id: str, name, value, description) -> dict[str, Any]:
    if name is not None and not validate_env_var_name(name, kind.noun):
        raise typer.Exit(1)
    payload = {k: v for k, v in (("name", name), ("value", value), ("description", description)) if v is not None}
    return client.patch(f"/environmentshub/{env_id}/{kind.path}/{item_id}", json=payload).get("data", {})


# ------------------------------------------------------------------------------------------------------- secrets
@secret_app.command("list")
@handle_errors
def env_secret_list(environment: Optional[str] = ENV_ARG, output: str = OUTPUT_OPT) -> None:
    """List all secrets for an environment."""
    _list(SECRET, environment, output)


@secret_app.command("create")
@handle_errors
def env_secret_create(environment: Optional[str] = ENV_ARG,
                      name: Optional[str] = typer.Option(None, "--name", "-n", help="Secret name (UPPER_SNAKE_CASE)"),
                      value: Optional[str] = typer.Option(None, "--value", "-v", help="Secret value"),
                      description: Optional[str] = typer.Option(None, "--description", "-d", help="Secret description"),
                      output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """Create an environment-specific secret."""
    _create(SECRET, environment, name, value, description, output)


@secret_app.command("update")
@handle_errors
def env_secret_update(environment: Optional[str] = ENV_ARG,
                      secret_id: Optional[str] = typer.Option(None, "--id", help="Secret ID (interactive selection if omitted)"),
                      name: Optional[str] = typer.Option(None, "--name", "-n", help="New secret name"),
                      value: Optional[str] = typer.Option(None, "--value", "-v", help="New secret value"),
                      description: Optional[str] = typer.Option(None, "--description", "-d", help="New description"),
                      output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """Update an environment-specific secret."""
    validate_output_format(output, console)
    owner, env_name = resolve_environment(environment)
    client = api()
    env_id = get_environment_id(client, owner, env_name)
    if not secret_id:
        secret_id = require_selection(fetch_items(client, env_id, SECRET), "update", f"No secrets to update for {owner}/{env_name}.").get("id")
    if not any_provided(name, value, description):
        console.print("\n[bold]What would you like to update?[/bold]")
        value = prompt_for_value("New value", required=False, hide_input=True) or None
        if not value:
            console.print("\n[dim]No changes made.[/dim]")
            raise typer.Exit()
    item = _patch(SECRET, client, env_id, secret_id, name, value, description)
    if output == "json":
        output_data_as_json(item, console)
        return
    console.print(f"[green]✓ Updated secret '{item.get('name')}' for {owner}/{env_name}[/green]")


@secret_app.command("delete")
@handle_errors
def env_secret_delete(environment: Optional[str] = ENV_ARG,
                      secret_id: Optional[str] = typer.Option(None, "--id", help="Secret ID (interactive selection if omitted)"),
                      yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation prompt")) -> None:  # fmt: skip
    """Delete an environment-specific secret."""
    owner, env_name = resolve_environment(environment)
    client = api()
    env_id = get_environment_id(client, owner, env_name)
    secrets = fetch_items(client, env_id, SECRET)
    if secret_id:
        label = next((s.get("name") for s in secrets if s.get("id") == secret_id), secret_id)
    else:
        chosen = require_selection(secrets, "delete", f"No secrets to delete for {owner}/{env_name}.")
        secret_id, label = chosen.get("id"), chosen.get("name")
    if not confirm_or_skip(f"Delete secret '{label}' from {owner}/{env_name}?", yes):
        raise _cancelled()
    client.delete(f"/environmentshub/{env_id}/secrets/{secret_id}")
    console.print(f"[green]✓ Deleted secret '{label}' from {owner}/{env_name}[/green]")


@secret_app.command("link")
@handle_errors
def env_secret_link(global_secret_id: str = typer.Argument(..., help="Global secret ID to link"),
                    environment: Optional[str] = ENV_ARG, output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """Link a global secret to an environment."""
    validate_output_format(output, console)
    owner, env_name = resolve_environment(environment)
    client = api()
    env_id = get_environment_id(client, owner, env_name)
    linked = client.post(f"/environmentshub/{env_id}/secrets/link/{global_secret_id}", json={}).get("data", {})
    if output == "json":
        output_data_as_json(linked, console)
        return
    console.print(f"[green]✓ Linked global secret '{linked.get('secretName', global_secret_id)}' to {owner}/{env_name}[/green]")


@secret_app.command("unlink")
@handle_errors
def env_secret_unlink(global_secret_id: str = typer.Argument(..., help="Global secret ID to unlink"),
                      environment: Optional[str] = ENV_ARG,
                      yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation prompt")) -> None:  # fmt: skip
    """Unlink a global secret from an environment."""
    owner, env_name = resolve_environment(environment)
    if not confirm_or_skip(f"Unlink global secret {global_secret_id} from {owner}/{env_name}?", yes):
        raise _cancelled()
    client = api()
    client.delete(f"/environmentshub/{get_environment_id(client, owner, env_name)}/secrets/link/{global_secret_id}")
    console.print(f"[green]✓ Unlinked global secret from {owner}/{env_name}[/green]")


# ------------------------------------------------------------------------------------------------------- variables
@var_app.command("list")
@handle_errors
def var_list(environment: Optional[str] = ENV_ARG, output: str = OUTPUT_OPT) -> None:
    """List all variables for an environment."""
    _list(VARIABLE, environment, output)


@var_app.command("create")
@handle_errors
def var_create(environment: Optional[str] = ENV_ARG,
               name: Optional[str] = typer.Option(None, "--name", "-n", help="Variable name (UPPER_SNAKE_CASE)"),
               value: Optional[str] = typer.Option(None, "--value", "-v", help="Variable value"),
               description: Optional[str] = typer.Option(None, "--description", "-d", help="Variable description"),
               output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """Create an environment variable."""
    _create(VARIABLE, environment, name, value, description, output)


@var_app.command("update")
@handle_errors
def var_update(var_id: str = typer.Argument(..., help="Variable ID to update"), environment: Optional[str] = ENV_ARG,
               name: Optional[str] = typer.Option(None, "--name", "-n", help="New variable name"),
               value: Optional[str] = typer.Option(None, "--value", "-v", help="New variable value"),
               description: Optional[str] = typer.Option(None, "--description", "-d", help="New description"),
               output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """Update an environment variable."""
    validate_output_format(output, console)
    if not any_provided(name, value, description):
        raise fail("At least one of --name, --value, or --description is required")
    owner, env_name = resolve_environment(environment)
    client = api()
    item = _patch(VARIABLE, client, get_environment_id(client, owner, env_name), var_id, name, value, description)
    if output == "json":
        output_data_as_json(item, console)
        return
    console.print(f"[green]✓ Updated variable '{item.get('name')}'[/green]")


@var_app.command("delete")
@handle_errors
def var_delete(var_id: str = typer.Argument(..., help="Variable ID to delete"), environment: Optional[str] = ENV_ARG,
               yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation prompt")) -> None:  # fmt: skip
    """Delete an environment variable."""
    owner, env_name = resolve_environment(environment)
    if not confirm_or_skip(f"Delete variable {var_id} from {owner}/{env_name}?", yes):
        raise _cancelled()
    client = api()
    client.delete(f"/environmentshub/{get_environment_id(client, owner, env_name)}/variables/{var_id}")
    console.print("[green]✓ Variable deleted[/green]")
