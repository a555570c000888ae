"""``prime gepa [run] ENV_OR_CONFIG …`` — GEPA prompt optimisation (reference: packages/prime/src/prime_cli/commands/gepa.py:19-59)."""

from __future__ import annotations

import typer

from ..verifiers_bridge import print_gepa_run_help, run_gepa_passthrough
from ._common import make_app
from ._passthrough import RAW_ARGS, require_leading_argument, wants_help

app = make_app("Run GEPA prompt optimization.", default_cmd="run")


@app.command("run", no_args_is_help=True, context_settings={**RAW_ARGS, "help_option_names": []})
def run_gepa_cmd(ctx: typer.Context, environment_or_config: str | None = typer.Argument(None, help="Environment name/slug or TOML config path")) -> None:
    """Run optimization with local-first environment resolution."""
    argv = list(ctx.args)
    wants_help(environment_or_config, argv, print_gepa_run_help)
    target = require_leading_argument(environment_or_config, "ENV_OR_CONFIG", "prime gepa run wordle --max-calls 100")
    run_gepa_passthrough(target, argv)
