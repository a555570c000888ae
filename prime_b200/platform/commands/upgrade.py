"""``prime upgrade`` — work out which tool installed the CLI (uv tool / pipx / pip) from where the interpreter lives and
run that tool's upgrade (behaviour: reference packages/prime/src/prime_cli/commands/upgrade.py:15-127).

Table-driven: ``INSTALLERS`` says how each kind is recognised and which commands to try, ``decide`` is the pure
version/flag logic (unit-testable without PyPI), the command glues them together.
"""

from __future__ import annotations

import shutil
import subprocess
import sys
from typing import Callable, NamedTuple

import typer
from packaging.version import Version

from .. import __version__
from ..utils.version_check import get_latest_pypi_version
from ._common import console, make_app

app = make_app("Upgrade the CLI to the latest version", invoke_without_command=True)
PACKAGE = "prime"


class Installer(NamedTuple):
    key: str
    path_marker: str | None  # substring of sys.executable that identifies it; None = fallback
    attempts: tuple[tuple[str, ...], ...]  # tried in order; first that exists AND exits 0 wins


INSTALLERS = (
    Installer("uv_tool", "uv/tools", (("uv", "tool", "upgrade", PACKAGE),)),
    Installer("pipx", "pipx/venvs", (("pipx", "upgrade", PACKAGE),)),
    Installer("pip", None, (("uv", "pip", "install", "--upgrade", PACKAGE), ("pip", "install", "--upgrade", PACKAGE))),
)
RECIPES = {i.key: [list(a) for a in i.attempts] for i in INSTALLERS}


def detect_install_method(executable: str | None = None) -> str:
    where = (executable or sys.executable).replace("\\", "/")
    return next(i.key for i in INSTALLERS if i.path_marker is None or i.path_marker in where)


def _attempt(cmd: list[str], runner: Callable[..., subprocess.CompletedProcess]) -> str | None:
    """Run one upgrade command; ``None`` on success, otherwise the (already rich-marked-up) reason it did not work."""
    try:
        done = runner(cmd, capture_output=True, text=True, timeout=120)
    except subprocess.TimeoutExpired:
        return "[red]Upgrade command timed out[/red]"
    except Exception as e:
        return f"[red]Error running upgrade: {e}[/red]"
    return None if done.returncode == 0 else f"[yellow]Command failed: {(done.stderr or '').strip()}[/yellow]"


def run_upgrade(method: str, runner=subprocess.run, which=shutil.which) -> bool:
    for cmd in RECIPES.get(method) or RECIPES["pip"]:
        if which(cmd[0]) is None:
            continue
        console.print(f"[dim]Running: {' '.join(cmd)}[/dim]")
        problem = _attempt(cmd, runner)
        if problem is None:
            return True
        console.print(problem)
    return False


def decide(installed: str, latest: str, *, check: bool, force: bool) -> tuple[str, bool]:
    """→ (``"current"`` | ``"report"`` | ``"upgrade"``, is a newer release out)."""
    newer = Version(installed) < Version(latest)
    if not newer and not force:
        return "current", newer
    return ("report" if check else "upgrade"), newer


MANUAL = "\n".join(f"  [dim]{' '.join(i.attempts[-1])}[/dim]" for i in INSTALLERS)


@app.callback(invoke_without_command=True)
def upgrade(ctx: typer.Context, check: bool = typer.Option(False, "--check", "-c", help="Only check for updates"),
            force: bool = typer.Option(False, "--force", "-f", help="Upgrade even when already current")) -> None:  # fmt: skip
    """Upgrade to the latest release."""
    if ctx.invoked_subcommand is not None:
        return
    latest = get_latest_pypi_version()
    if latest is None:
        console.print("[red]Could not fetch latest version from PyPI[/red]")
        raise typer.Exit(1)
    console.print(f"[cyan]Installed version:[/cyan] {__version__}\n[cyan]Latest version:[/cyan]    {latest}")
    action, newer = decide(__version__, latest, check=check, force=force)
    if action == "current":
        console.print("\n[green]✓ You are already on the latest version![/green]")
        return
    if newer:
        console.print(f"\n[yellow]A newer version is available: {latest}[/yellow]")
    if action == "report":
        if newer:
            console.print("\n[dim]Run 'prime upgrade' to upgrade[/dim]")
        return
    method = detect_install_method()
    console.print(f"\n[dim]Detected install method: {method}[/dim]")
    if not run_upgrade(method):
        console.print(f"\n[red]Upgrade failed. Try manually:[/red]\n{MANUAL}")
        raise typer.Exit(1)
    console.print(f"\n[green]✓ Successfully upgraded to {latest}![/green]")
