"""Commands that hand their whole argv to a ``verifiers`` module (``lab setup``, ``gepa run``): one builder, so that the
per-command files only declare what differs — the help printer, the target and the positional argument's name."""

from __future__ import annotations

import subprocess
from typing import Callable

import typer

from ..verifiers_bridge import is_help_request
from ._common import console

RAW_ARGS = {"allow_extra_args": True, "ignore_unknown_options": True}


def wants_help(first: str | None, rest: list[str], show: Callable[[], None]) -> None:
    """``-h/--help`` anywhere in a pass-through argv prints OUR help for the wrapped module and exits 0."""
    if is_help_request(first or "", rest):
        show()
        raise typer.Exit(0)


def require_leading_argument(value: str | None, placeholder: str, example: str) -> str:
    """The wrapped tools take their target first; say so instead of letting argparse of the child complain."""
    problem = None
    if value is None:
        problem = f"Missing argument '{placeholder}'."
    elif value.startswith("-"):
        problem = "Environment/config must be the first argument."
    if problem:
        console.print(f"[red]Error:[/red] {problem}\n[dim]Example: {example}[/dim]")
        raise typer.Exit(2)
    return value  # type: ignore[return-value]


def run_module(command: list[str]) -> None:
    """Run the child in the foreground; propagate its exit status."""
    status = subprocess.run(command).returncode
    if status:
        raise typer.Exit(status)
