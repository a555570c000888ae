"""Locate the interpreter and module names used to drive the external ``verifiers`` toolkit.

Contract (reference: packages/prime/src/prime_cli/verifiers_plugin.py:46-146): interpreter search order
UV_PROJECT_ENVIRONMENT → VIRTUAL_ENV → nearest ancestor ``.venv`` next to a ``pyproject.toml`` → ``sys.executable``,
accepting a candidate only if it can import the verifiers CLI; module names come from
``verifiers.cli.plugins.prime.get_plugin()`` when available (API version 1), else built-in defaults.
"""

from __future__ import annotations

import importlib
import os
import subprocess
import sys
from dataclasses import dataclass, fields
from functools import lru_cache
from pathlib import Path
from typing import Iterator, Sequence

from .utils.plain import get_console

EXPECTED_PLUGIN_API_VERSION = 1
PROBE_MODULE = "verifiers.cli.commands.eval"
_PROBE = "import importlib.util, sys; raise SystemExit(0 if importlib.util.find_spec(sys.argv[1]) else 1)"


def venv_python(root: Path) -> Path:
    return root / ("Scripts/python.exe" if os.name == "nt" else "bin/python")


@lru_cache(maxsize=32)
def can_import(python: str, module: str, cwd: str) -> bool:
    try:
        return subprocess.run([python, "-c", _PROBE, module], cwd=cwd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode == 0
    except Exception:
        return False


def candidate_interpreters(workspace: Path) -> Iterator[Path]:
    for var in ("UV_PROJECT_ENVIRONMENT", "VIRTUAL_ENV"):
        if os.environ.get(var):
            yield venv_python(Path(os.environ[var]))
    for d in (workspace, *workspace.parents):
        if (d / "pyproject.toml").is_file():
            yield venv_python(d / ".venv")


def resolve_workspace_python(cwd: Path | None = None) -> str:
    ws = (cwd or Path.cwd()).resolve()
    for cand in candidate_interpreters(ws):
        if cand.exists() and can_import(str(cand), PROBE_MODULE, str(ws)):
            return str(cand)
    return sys.executable


@dataclass(frozen=True)
class PrimeVerifiersPlugin:
    api_version: int = EXPECTED_PLUGIN_API_VERSION
    eval_module: str = "verifiers.cli.commands.eval"
    gepa_module: str = "verifiers.cli.commands.gepa"
    install_module: str = "verifiers.cli.commands.install"
    init_module: str = "verifiers.cli.commands.init"
    setup_module: str = "verifiers.cli.commands.setup"
    build_module: str = "verifiers.cli.commands.build"
    tui_module: str = "verifiers.cli.tui"

    def build_module_command(self, module_name: str, args: Sequence[str] | None = None) -> list[str]:
        return [resolve_workspace_python(), "-m", module_name, *(args or ())]


def load_verifiers_prime_plugin(console=None) -> PrimeVerifiersPlugin:
    sink = console or get_console(stderr=True)

    def fallback(why: str) -> PrimeVerifiersPlugin:
        sink.print(f"[yellow]Warning:[/yellow] {why}. Falling back to built-in command mapping.")
        return PrimeVerifiersPlugin()

    try:
        module = importlib.import_module("verifiers.cli.plugins.prime")
    except Exception as e:
        return fallback(f"Could not import verifiers prime plugin ({e})")
    getter = getattr(module, "get_plugin", None)
    if not callable(getter):
        return fallback("verifiers prime plugin module does not expose a callable get_plugin()")
    try:
        ext = getter()
    except Exception as e:
        return fallback(f"Failed to load verifiers plugin ({e})")
    version = getattr(ext, "api_version", None)
    if version != EXPECTED_PLUGIN_API_VERSION:
        sink.print(f"[yellow]Warning:[/yellow] verifiers plugin API version mismatch (got {version}, expected "
                   f"{EXPECTED_PLUGIN_API_VERSION}). Continuing with compatibility behavior.")  # fmt: skip
    defaults = PrimeVerifiersPlugin()
    names = {f.name: getattr(ext, f.name, getattr(defaults, f.name)) for f in fields(PrimeVerifiersPlugin) if f.name != "api_version"}
    return PrimeVerifiersPlugin(api_version=int(version or EXPECTED_PLUGIN_API_VERSION), **names)
