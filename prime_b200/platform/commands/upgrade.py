"""``prime upgrade`` — detect uv-tool / pipx / pip installs and run the matching upgrade
(reference: packages/prime/src/prime_cli/commands/upgrade.py:15-127)."""

from __future__ import annotations

import shutil
import subprocess
import sys

import typer
from packaging.version import Version

from .. import __version__
from ..utils.version_check import get_latest_pypi_version
from ._common import console, make_app

app = make_app("Upgrade the CLI to the latest version", invoke_without_command=True)
PACKAGE = "prime"
RECIPES: dict[str, list[list[str]]] = {
    "uv_tool": [["uv", "tool", "upgrade", PACKAGE]],
    "pipx": [["pipx", "upgrade", PACKAGE]],
    "pip": [["uv", "pip", "install", "--upgrade", PACKAGE], ["pip", "install", "--upgrade", PACKAGE]],
}


def detect_install_method(executable: str | None = None) -> str:
    exe = (executable or sys.executable).replace("\\", "/")
    if "uv/tools" in exe:
        return "uv_tool"
    if "pipx/venvs" in exe:
        return "pipx"
    return "pip"


def run_upgrade(method: str, runner=subprocess.run, which=shutil.which) -> bool:
    for cmd in RECIPES.get(method, RECIPES["pip"]):
        if which(cmd[0]) is None:
            continue
        console.print(f"[dim]Running: {' '.join(cmd)}[/dim]")
        try:
            r = runner(cmd, capture_output=True, text=True, timeout=120)
        except subprocess.TimeoutExpired:
            console.print("[red]Upgrade command timed out[/red]")
            continue
        except Exception as e:
            console.print(f"[red]Error running upgrade: {e}[/red]")
            continue
        if r.returncode == 0:
            return True
        console.print(f"[yellow]Command failed: {(r.stderr or '').strip()}[/yellow]")
    return False


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
    newer = Version(__version__) < Version(latest)
    if not newer and not force:
        console.print("\n[green]✓ You are already on the latest version![/green]")
        raise typer.Exit(0)
    if newer:
        console.print(f"\n[yellow]A newer version is available: {latest}[/yellow]")
    if check:
        if newer:
            console.print("\n[dim]Run 'prime upgrade' to upgrade[/dim]")
        raise typer.Exit(0)
    method = detect_install_method()
    console.print(f"\n[dim]Detected install method: {method}[/dim]")
    if run_upgrade(method):
        console.print(f"\n[green]✓ Successfully upgraded to {latest}![/green]")
        return
    console.print("\n[red]Upgrade failed. Try manually:[/red]\n  [dim]uv tool upgrade prime[/dim]\n  [dim]pipx upgrade prime[/dim]\n  [dim]pip install --upgrade prime[/dim]")
    raise typer.Exit(1)
