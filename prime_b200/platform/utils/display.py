"""Table / JSON output helpers and status colour maps
(reference: packages/prime/src/prime_cli/utils/display.py:13-93)."""

from __future__ import annotations

import json
from typing import Any, Iterable, Sequence

from rich.table import Table

from .plain import get_console, is_plain_mode

POD_STATUS_COLORS = {"ACTIVE": "green", "RUNNING": "green", "PROVISIONING": "yellow", "PENDING": "yellow",
                     "INSTALLING": "yellow", "STOPPED": "dim", "TERMINATED": "red", "ERROR": "red", "FAILED": "red"}  # fmt: skip
SANDBOX_STATUS_COLORS = {"RUNNING": "green", "PENDING": "yellow", "PROVISIONING": "yellow", "STOPPED": "dim",
                         "ERROR": "red", "TERMINATED": "red", "TIMEOUT": "red"}  # fmt: skip
RUN_STATUS_COLORS = {"QUEUED": "yellow", "PENDING": "yellow", "RUNNING": "green", "COMPLETED": "cyan",
                     "STOPPED": "dim", "FAILED": "red", "CANCELLED": "dim"}  # fmt: skip
DEPLOYMENT_STATUS_COLORS = {"DEPLOYED": "green", "DEPLOYING": "yellow", "UNLOADING": "yellow",
                            "NOT_DEPLOYED": "dim", "DEPLOY_FAILED": "red", "UNLOAD_FAILED": "red"}  # fmt: skip


def status_color(status: str | None, mapping: dict[str, str], default: str = "white") -> str:
    return mapping.get((status or "").upper(), default)


def colorize(status: str | None, table: dict[str, str]) -> str:
    s = status or "UNKNOWN"
    return f"[{status_color(s, table)}]{s}[/]"


def build_table(title: str | None, columns: Sequence[str | tuple[str, str]], rows: Iterable[Sequence[Any]] = (),
                show_lines: bool = False) -> Table:  # fmt: skip
    t = Table(title=title, show_lines=show_lines)
    for c in columns:
        if isinstance(c, tuple):
            t.add_column(c[0], style=c[1])
        else:
            t.add_column(c)
    blank = "-" if is_plain_mode() else ""  # plain tables are split on whitespace by their readers: an empty cell would shift the columns
    for r in rows:
        t.add_row(*[blank if v is None or v == "" else str(v) for v in r])
    return t


def output_data_as_json(data: Any, console=None) -> None:
    (console or get_console()).file.write(json.dumps(data, indent=2, default=str) + "\n")


def validate_output_format(output: str, console=None) -> str:
    fmt = (output or "table").lower()
    if fmt not in ("table", "json"):
        import typer

        (console or get_console()).print(f"[red]Error:[/red] Invalid output format {fmt!r}: --output takes 'table' or 'json'")
        raise typer.Exit(1)
    return fmt


def get_eval_viewer_url(evaluation_id: str) -> str:
    """Dashboard page of one evaluation (reference: packages/prime/src/prime_cli/utils/display.py:47-50)."""
    from ..core import Config

    return f"{Config(writable=False).frontend_url.rstrip('/')}/dashboard/evaluations/{evaluation_id}"

