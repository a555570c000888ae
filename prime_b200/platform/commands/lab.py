"""``prime lab setup`` — bootstrap a verifiers training workspace (reference: packages/prime/src/prime_cli/commands/lab.py:15-33)."""

from __future__ import annotations

import typer

from ..verifiers_bridge import print_lab_setup_help
from ..verifiers_plugin import load_verifiers_prime_plugin
from ._common import console, make_app
from ._passthrough import RAW_ARGS, run_module, wants_help

app = make_app("Lab commands for verifiers development")


@app.command(add_help_option=False, context_settings=RAW_ARGS)
def setup(ctx: typer.Context) -> None:
    """Set up a verifiers training workspace."""
    argv = list(ctx.args)
    wants_help(None, argv, print_lab_setup_help)
    plugin = load_verifiers_prime_plugin(console=console)
    run_module(plugin.build_module_command(plugin.setup_module, argv))
