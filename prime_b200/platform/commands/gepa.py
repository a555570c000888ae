"""``prime gepa [run] ENV_OR_CONFIG …`` — GEPA prompt optimisation pass-through
(reference: packages/prime/src/prime_cli/commands/gepa.py:19-59)."""

from __future__ import annotations

import typer

from ..verifiers_bridge import is_help_request, print_gepa_run_help, run_gepa_passthrough
from ._common import console, make_app

app = make_app("Run GEPA prompt optimization.", default_cmd="run")
_EXAMPLE = "[dim]Example: prime gepa run wordle --max-calls 100[/dim]"


@app.command("run", no_args_is_help=True,
             context_settings={"allow_extra_args": True, "ignore_unknown_options": True, "help_option_names": []})  # fmt: skip
def run_gepa_cmd(ctx: typer.Context, environment_or_config: str | None = typer.Argument(None, help="Environment name/slug or TOML config path")) -> None:
    """Run optimization with local-first environment resolution."""
    args = list(ctx.args)
    if is_help_request(environment_or_config or "", args):
        print_gepa_run_help()
        raise typer.Exit(0)
    if environment_or_config is None:
        console.print(f"[red]Error:[/red] Missing argument 'ENV_OR_CONFIG'.\n{_EXAMPLE}")
        raise typer.Exit(2)
    if environment_or_config.startswith("-"):
        console.print(f"[red]Error:[/red] Environment/config must be the first argument.\n{_EXAMPLE}")
        raise typer.Exit(2)
    run_gepa_passthrough(environment_or_config, args)
