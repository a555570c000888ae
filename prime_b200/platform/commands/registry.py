"""``prime registry {list,check-image}`` (reference: packages/prime/src/prime_cli/commands/registry.py:53-142)."""

from __future__ import annotations

from typing import Optional

import typer

from ..sandboxes import TemplateClient
from ..utils.json_help import list_json_help
from ..utils.time_utils import human_age, iso_timestamp
from ._common import OUTPUT_OPT, api, console, emit, handle_errors, make_app

app = make_app("Private registry credentials and image checks")


@app.command("list", epilog=list_json_help("credentials", {"id": "str", "name": "str", "server": "str", "scope": "str", "team_id": "str|null", "user_id": "str|null",
                                                          "created_at": "str", "updated_at": "str", "age": "str"}))
@handle_errors
def list_registry_credentials(output: str = OUTPUT_OPT) -> None:
    """List registry credentials usable by sandboxes."""
    creds = TemplateClient(api()).list_registry_credentials()
    # scope = the owning team's id, else "user:<id>", else "personal"; an empty server means Docker Hub (the row scripts have always parsed)
    rows = [{"id": c.id, "name": c.name, "server": c.server or "registry-1.docker.io",
             "scope": c.team_id or (f"user:{c.user_id}" if c.user_id else "personal"), "team_id": c.team_id, "user_id": c.user_id,
             "created_at": iso_timestamp(c.created_at), "updated_at": iso_timestamp(c.updated_at), "age": human_age(c.created_at)} for c in creds]  # fmt: skip
    emit(output, {"credentials": rows, "total_count": len(rows)}, "Registry Credentials",
         [("ID", "cyan"), ("Name", "green"), "Server", "Scope", ("Created", "magenta")],
         [[r["id"], r["name"], r["server"], r["scope"], f"{r['created_at']} ({r['age']})"] for r in rows])  # fmt: skip


@app.command("check-image")
@handle_errors
def check_docker_image(image: str = typer.Argument(..., help="Image reference, e.g. ghcr.io/org/app:tag"),
                       registry_credentials_id: Optional[str] = typer.Option(None, "--registry-credentials-id", help="Credentials for private images")) -> None:  # fmt: skip
    """Check that the platform can pull an image."""
    r = TemplateClient(api()).check_docker_image(image, registry_credentials_id)
    console.print(f"[green]✓ Accessible[/green] {r.details}" if r.accessible else f"[red]✗ Not accessible[/red] {r.details}")
    if not r.accessible:
        raise typer.Exit(1)
