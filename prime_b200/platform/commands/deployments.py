"""``prime deployments {list,create,delete}`` — serve / unload trained LoRA adapters, with state-machine guards on
``deployment_status`` (reference: packages/prime/src/prime_cli/commands/deployments.py:34-292)."""

from __future__ import annotations

from typing import Optional

import typer

from ..api.deployments import DEPLOYABLE_FROM, UNLOADABLE_FROM, Adapter, DeploymentsClient
from ..core import APIError, Config
from ..utils.display import DEPLOYMENT_STATUS_COLORS, colorize
from ..utils.json_help import list_json_help
from ..utils.plain import get_console
from ..utils.time_utils import format_time_ago, parse_dt
from ._common import OUTPUT_OPT, api, console, emit, fail, handle_errors, make_app, paginate_hint

app = make_app("Deploy trained adapters for inference")
IN_FLIGHT = {"DEPLOYING", "UNLOADING"}


def deploy_blocker(a: Adapter) -> tuple[str, int] | None:
    """(message, exit code) when the adapter cannot be deployed right now; None when it can."""
    if a.status != "READY":
        return (f"Model is not ready for deployment (status {a.status}); only READY models can be deployed.", 1)
    if a.deployment_status == "DEPLOYED":
        return ("Model is already deployed.", 0)
    if a.deployment_status in IN_FLIGHT:
        return (f"Model deployment is in progress ({a.deployment_status}).", 1)
    if a.deployment_status not in DEPLOYABLE_FROM:
        return (f"Cannot deploy model in state {a.deployment_status}.", 1)
    return None


def unload_blocker(a: Adapter) -> tuple[str, int] | None:
    if a.deployment_status == "NOT_DEPLOYED":
        return ("Model is not deployed.", 0)
    if a.deployment_status in IN_FLIGHT:
        return (f"Model deployment is in progress ({a.deployment_status}).", 1)
    if a.deployment_status not in UNLOADABLE_FROM:
        return (f"Cannot unload model in state {a.deployment_status}.", 1)
    return None


def _stop(blocker: tuple[str, int]) -> None:
    msg, code = blocker
    console.print(f"[yellow]{msg}[/yellow]" if code == 0 else f"[red]Error:[/red] {msg}")
    raise typer.Exit(code)


@app.command("list", epilog=list_json_help("models", {"id": "str", "display_name": "str|null", "base_model": "str", "status": "str", "deployment_status": "str"}))
@handle_errors
def list_deployments(team: Optional[str] = typer.Option(None, "--team", "-t", help="Filter by team ID"),
                     num: int = typer.Option(20, "--num", "-n", help="Items per page"),
                     page: int = typer.Option(1, "--page", "-p", help="Page number"), output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """List trained adapters and where they stand (reference flags: packages/prime/src/prime_cli/commands/deployments.py:34-40)."""
    if num < 1 or page < 1:
        raise fail("--num and --page must be at least 1")
    limit, offset = num, (page - 1) * num
    client = DeploymentsClient(api())
    adapters, total = client.list_adapters(team_id=team or Config(writable=False).team_id, limit=limit, offset=offset)
    try:  # which base models can currently be served: every row says whether `deployments create` would be accepted
        deployable: list[str] | None = client.get_deployable_models()
    except Exception:  # noqa: BLE001 — the listing is still useful without the column
        get_console(stderr=True).print("[dim]Warning: Could not fetch deployable models list.[/dim]")  # stderr: stdout may be JSON
        deployable = None
    models = []
    for a in adapters:
        row = a.model_dump(mode="json")
        for key in ("created_at", "updated_at", "deployed_at"):  # "2025-01-01 00:00:00+00:00": the timestamp form this command has always printed
            if row.get(key):
                row[key] = str(parse_dt(row[key]))
        if deployable is not None:
            row["deployable"] = a.base_model in deployable
        models.append(row)
    mark = lambda a: "[dim]-[/dim]" if deployable is None else ("[green]Yes[/green]" if a.base_model in deployable else "[red]No[/red]")  # noqa: E731
    emit(output, {"models": models, "total": total, "page": page, "per_page": num}, f"Models (Total: {total})",
         [("ID", "cyan"), "Name", ("Base model", "blue"), "Step", "Status", "Deployment", "Deployable", ("Created", "magenta")],
         [[a.id, a.display_name or "", a.base_model, a.step if a.step is not None else "", a.status,
           colorize(a.deployment_status, DEPLOYMENT_STATUS_COLORS), mark(a), format_time_ago(a.created_at)] for a in adapters],
         paginate_hint(total, offset, limit, "models"))  # fmt: skip


@app.command("create")
@handle_errors
def create_deployment(ctx: typer.Context, model_id: Optional[str] = typer.Argument(None, help="Model (adapter) ID"),
                      yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation")) -> None:  # fmt: skip
    """Make a READY adapter available for inference."""
    if model_id is None:
        console.print(ctx.get_help())
        raise typer.Exit(0)
    client = DeploymentsClient(api())
    a = client.get_adapter(model_id)
    if (blocker := deploy_blocker(a)) is not None:
        _stop(blocker)
    try:
        if a.base_model not in client.get_deployable_models():
            console.print(f"[red]Error:[/red] Base model [yellow]{a.base_model}[/yellow] is not currently available for LoRA deployment.")
            raise typer.Exit(1)
    except APIError:
        console.print("[dim]Warning: could not verify base-model deployability; proceeding.[/dim]")
    console.print(f"[bold]Deploying model:[/bold]\n  ID: {a.id}" + (f"\n  Name: {a.display_name}" if a.display_name else "") + f"\n  Base Model: {a.base_model}\n")
    if not yes and not typer.confirm("Are you sure you want to deploy this model?"):
        console.print("Cancelled.")
        raise typer.Exit(0)
    updated = client.deploy_adapter(model_id)
    console.print(f"[green]Deployment initiated.[/green] Status: [yellow]{updated.deployment_status}[/yellow]")
    console.print("[dim]Check progress with 'prime deployments list'. Once deployed, call it as model "
                  f"'{a.base_model}:{a.id}' on {Config(writable=False).inference_url}/chat/completions[/dim]")  # fmt: skip


@app.command("delete")
@handle_errors
def delete_deployment(ctx: typer.Context, model_id: Optional[str] = typer.Argument(None, help="Model (adapter) ID")) -> None:
    """Unload an adapter from serving (its files stay stored)."""
    if model_id is None:
        console.print(ctx.get_help())
        raise typer.Exit(0)
    client = DeploymentsClient(api())
    if (blocker := unload_blocker(client.get_adapter(model_id))) is not None:
        _stop(blocker)
    updated = client.unload_adapter(model_id)
    console.print(f"[green]Unload initiated.[/green] Status: [yellow]{updated.deployment_status}[/yellow]")
