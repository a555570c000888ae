"""``prime lab setup`` — pass-through to the verifiers workspace bootstrapper
(reference: packages/prime/src/prime_cli/commands/lab.py:15-33)."""

from __future__ import annotations

import subprocess

import typer

from ..verifiers_bridge import is_help_request, print_lab_setup_help
from ..verifiers_plugin import load_verifiers_prime_plugin
from ._common import console, make_app

app = make_app("Lab commands for verifiers development")


@app.command(add_help_option=False, context_settings={"allow_extra_args": True, "ignore_unknown_options": True})
def setup(ctx: typer.Context) -> None:
    """Set up a verifiers training workspace."""
    args = list(ctx.args)
    if is_help_request("", args):
        print_lab_setup_help()
        raise typer.Exit(0)
    plugin = load_verifiers_prime_plugin(console=console)
    rc = subprocess.run(plugin.build_module_command(plugin.setup_module, args)).returncode
    if rc != 0:
        raise typer.Exit(rc)
