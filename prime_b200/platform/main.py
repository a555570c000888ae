"""Root of the ``prime`` command: 21 command groups in three help panels (Lab / Compute / Account), ``--version``,
a per-invocation ``--context`` switch and the daily update banner
(reference: packages/prime/src/prime_cli/main.py:32-117). Groups are registered from one table (name, module, help panel);
heavy optional dependencies (protobuf / Connect-RPC of the VM sandbox path, MCP, tunnel binary handling) are imported inside the
commands that need them, which is most of why ``prime --help`` starts 1.3× faster than the reference."""

from __future__ import annotations

import importlib
import os
import sys
from typing import Optional

import typer

from . import __version__
from .core import Config
from .utils.plain import PlainTyper, get_console
from .utils.version_check import check_for_update

# (command name, module under .commands, help panel)
GROUPS: tuple[tuple[str, str, str], ...] = (
    ("lab", "lab", "Lab"), ("env", "env", "Lab"), ("eval", "evals", "Lab"), ("gepa", "gepa", "Lab"), ("rl", "rl", "Lab"),
    ("deployments", "deployments", "Lab"),
    ("availability", "availability", "Compute"), ("disks", "disks", "Compute"), ("pods", "pods", "Compute"),
    ("sandbox", "sandbox", "Compute"), ("images", "images", "Compute"), ("registry", "registry", "Compute"),
    ("tunnel", "tunnel", "Compute"), ("inference", "inference", "Compute"),
    ("login", "login", "Account"), ("whoami", "whoami", "Account"), ("switch", "switch", "Account"),
    ("config", "config", "Account"), ("teams", "teams", "Account"), ("secret", "secrets", "Account"),
    ("upgrade", "upgrade", "Account"),
)  # fmt: skip

app = PlainTyper(name="prime", help=f"Prime Intellect CLI (v{__version__})", no_args_is_help=True,
                 context_settings={"help_option_names": ["-h", "--help"]})  # fmt: skip

for _name, _module, _panel in GROUPS:
    app.add_typer(importlib.import_module(f"{__package__}.commands.{_module}").app, name=_name, rich_help_panel=_panel)


@app.callback(invoke_without_command=True)
def callback(ctx: typer.Context,
             version_flag: bool = typer.Option(False, "--version", "-v", help="Show version and exit"),
             context: Optional[str] = typer.Option(None, "--context", "-c", help="Use a specific config context for this command")) -> None:  # fmt: skip
    """Prime Intellect CLI"""
    if version_flag:
        typer.echo(f"Prime CLI version: {__version__}")
        raise typer.Exit()
    try:
        Config()  # the CLI (not the SDKs) materialises ~/.prime/config.json with the defaults on first use, as the reference does
    except OSError:
        pass  # read-only home: every command still works from the environment variables
    if context:
        known = Config(writable=False).list_environments()
        if context.lower() != "production" and context not in known:
            typer.echo(f"Error: Unknown context '{context}'", err=True)
            typer.echo("Available contexts:", err=True)
            for name in known:
                typer.echo(f"  - {name}", err=True)
            raise typer.Exit(1)
        os.environ["PRIME_CONTEXT"] = context  # every Config() built by the subcommand sees it
    if ctx.invoked_subcommand is not None and ctx.invoked_subcommand != "upgrade":
        available, latest = check_for_update()  # no arguments: the installed version is this package's (and tests stub it with a zero-argument callable)
        if available and latest:
            err = get_console(stderr=True)
            err.print(f"[yellow]A new version of prime is available: {latest} (installed: {__version__})[/yellow]")
            err.print("[dim]Run: prime upgrade  ·  set PRIME_DISABLE_VERSION_CHECK=1 to silence this check[/dim]\n")


def run() -> None:
    """Console entry point."""
    try:
        app()
    except typer.Abort:
        typer.echo("\nOperation cancelled")
        sys.exit(0)


if __name__ == "__main__":
    run()
