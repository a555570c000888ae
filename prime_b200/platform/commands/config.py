"""``prime config …`` — view/set identity + endpoints, named contexts
(reference: packages/prime/src/prime_cli/commands/config.py:35-396)."""

from __future__ import annotations

import os
import re
from typing import Optional

import typer

from ..core import APIClient, Config
from ..utils.display import build_table
from ._common import console, make_app

app = make_app("Configure the CLI")
TEAM_ID = re.compile(r"c[a-z0-9]{24}\Z")  # CUID v1


def validate_team_id(team_id: str) -> bool:
    return not team_id or bool(TEAM_ID.match(team_id))


def mask(key: str, sep: str = "...") -> str:
    return f"{key[:6]}{sep}{key[-4:]}" if len(key) > 10 else "***"


def _from_env(*names: str) -> str:
    return " (from env var)" if any((os.getenv(n) or "").strip() for n in names) else ""


def _remember_user(config: Config, api_key: str) -> None:
    try:
        data = APIClient(api_key=api_key, config=config).get("/user/whoami").get("data")
        if isinstance(data, dict) and data.get("id"):
            config.set_user_id(data["id"])
            config.update_current_environment_file()
    except Exception:
        pass


@app.command()
def view() -> None:
    """Show the effective configuration (env vars override the file)."""
    c = Config()
    s = c.view()
    key = (mask(s["api_key"]) + _from_env("PRIME_API_KEY")) if s["api_key"] else "Not set"
    if s["team_id"]:
        team = f"{s['team_id']} (from env var)" if c.team_id_from_env else (f"{s['team_name']} ({s['team_id']})" if s["team_name"] else s["team_id"])
    else:
        team = "Personal Account"
    rows = [
        ("Current Environment", s["current_environment"]),
        ("API Key", key),
        ("Team", team),
        ("User ID", (s["user_id"] + _from_env("PRIME_USER_ID")) if s["user_id"] else "Not set"),
        ("Base URL", s["base_url"] + _from_env("PRIME_API_BASE_URL", "PRIME_BASE_URL")),
        ("Frontend URL", s["frontend_url"] + _from_env("PRIME_FRONTEND_URL")),
        ("Inference URL", s["inference_url"] + _from_env("PRIME_INFERENCE_URL")),
        ("SSH Key Path", s["ssh_key_path"] + _from_env("PRIME_SSH_KEY_PATH")),
        ("Share Resources With Team", str(s["share_resources_with_team"])),
    ]
    console.print(build_table("Prime CLI Configuration", [("Setting", "cyan"), ("Value", "green")], rows))


@app.command("set-api-key")
def set_api_key(api_key: Optional[str] = typer.Argument(None, help="API key; prompted securely when omitted")) -> None:
    """Store the API key (empty input clears it)."""
    if api_key is None:
        api_key = typer.prompt("Enter your Prime Intellect API key (or press Enter to clear)", hide_input=True, default="", show_default=False)
    c = Config()
    c.set_api_key(api_key)
    if not api_key:
        console.print("[green]API key cleared successfully![/green]")
        return
    _remember_user(c, api_key)
    console.print(f"[green]API key {mask(api_key, '***')} configured successfully![/green]")
    console.print("[blue]Verify with 'prime config view'[/blue]")
    console.print("\n[yellow]Tip: create keys at https://app.primeintellect.ai/dashboard/tokens[/yellow]")


@app.command("set-team-id")
def set_team_id(team_id: str = typer.Argument(..., help="Team ID (empty string = personal account)")) -> None:
    """Select a team by id."""
    if not validate_team_id(team_id):
        console.print("[red]Invalid team ID format. Expected a CUID like 'c' followed by 24 lowercase letters/digits.[/red]")
        raise typer.Exit(1)
    c = Config()
    name = role = None
    if team_id:
        try:  # best effort: enrich with name/role so `config view` is readable
            from .teams import fetch_teams

            for t in fetch_teams(APIClient(config=c)):
                if t.get("teamId") == team_id:
                    name, role = t.get("name"), t.get("role")
        except Exception:
            pass
    c.set_team(team_id or None, team_name=name, team_role=role)
    c.update_current_environment_file()
    if not team_id:
        console.print("[green]Team ID cleared. Using personal account.[/green]")
    elif name:
        console.print(f"[green]Team '{name}' ({team_id}) configured successfully![/green]")
    else:  # the listing was unreachable or does not contain it: the id is stored all the same
        console.print(f"[green]Team ID '{team_id}' configured successfully![/green]")


@app.command("remove-team-id")
def remove_team_id() -> None:
    """Back to the personal account."""
    c = Config()
    c.set_team(None)
    c.update_current_environment_file()
    console.print("[green]Team ID removed. Using personal account.[/green]")


def _url_setter(name: str, attr: str, prompt: str):
    def cmd(url: Optional[str] = typer.Argument(None, help=f"New {name} URL; prompted when omitted")) -> None:
        c = Config()
        if url is None:
            url = typer.prompt(prompt, default=getattr(c, attr))
        if not re.match(r"https?://", url or ""):
            console.print("[red]URL must start with http:// or https://[/red]")
            raise typer.Exit(1)
        getattr(c, f"set_{attr}")(url)
        c.update_current_environment_file()
        console.print(f"[green]{name} URL set to {getattr(c, attr)}[/green]")

    cmd.__doc__ = f"Set the {name} URL."
    return cmd


app.command("set-base-url")(_url_setter("API base", "base_url", "Enter the API base URL"))
app.command("set-frontend-url")(_url_setter("frontend", "frontend_url", "Enter the frontend URL"))
app.command("set-inference-url")(_url_setter("inference", "inference_url", "Enter the inference URL"))


@app.command("set-share-resources-with-team", no_args_is_help=True)
def set_share(enabled: bool = typer.Argument(..., help="true / false")) -> None:
    """New resources are shared with the active team by default."""
    c = Config()
    c.set_share_resources_with_team(enabled)
    c.update_current_environment_file()
    console.print(f"[green]Share resources with team: {enabled}[/green]")


@app.command("set-ssh-key-path", no_args_is_help=True)
def set_ssh_key_path(path: str = typer.Argument(..., help="Private key used by 'prime pods ssh'")) -> None:
    """Set the SSH private key path."""
    c = Config()
    c.set_ssh_key_path(path)
    console.print(f"[green]SSH key path set to {c.ssh_key_path}[/green]")


@app.command()
def reset(yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation")) -> None:
    """Restore defaults (clears the API key)."""
    if not yes and not typer.confirm("Reset all configuration to defaults?", default=False):
        raise typer.Exit()
    Config().reset()
    console.print("[green]Configuration reset.[/green]")


@app.command("use", no_args_is_help=True)
def use_environment(name: str = typer.Argument(..., help="Context name ('production' is built in)")) -> None:
    """Switch to a saved context."""
    if not Config().load_environment(name):
        console.print(f"[red]Unknown environment '{name}'.[/red] Saved: {', '.join(Config().list_environments())}")
        raise typer.Exit(1)
    console.print(f"[green]Now using '{Config().current_environment}'.[/green]")


@app.command("save", no_args_is_help=True)
def save_env(name: str = typer.Argument(..., help="Name for the context")) -> None:
    """Snapshot the current settings as a named context."""
    try:
        saved = Config().save_environment(name)
    except ValueError as e:
        console.print(f"[red]{e}[/red]")
        raise typer.Exit(1)
    console.print(f"[green]Saved context '{saved}'.[/green] Activate with: prime config use {saved}")


@app.command("envs")
def list_envs() -> None:
    """List saved contexts."""
    c = Config()
    for n in c.list_environments():
        console.print(f"  {n}" + (" [green](current)[/green]" if n == c.current_environment else ""))
