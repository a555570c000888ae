"""Shared plumbing for command modules: client construction, uniform API-error handling, table/JSON output."""

from __future__ import annotations

import functools
import sys
from typing import Any, Callable, Iterable, Sequence

import typer

from ..core import APIClient, APIError, Config, ValidationError
from ..utils.display import build_table, output_data_as_json, validate_output_format
from ..utils.plain import PlainTyper, default_group, get_console

console = get_console()
OUTPUT_OPT = typer.Option("table", "--output", "-o", help="Output format: table or json")


def make_app(help: str, default_cmd: str | None = None, **kw: Any) -> PlainTyper:
    if default_cmd:
        kw["cls"] = default_group(default_cmd)
    return PlainTyper(help=help, no_args_is_help=default_cmd is None and not kw.get("invoke_without_command"), **kw)


def api(require_auth: bool = True) -> APIClient:
    """The control-plane client of a command.  Injection point: a command module's own ``APIClient`` name — when a harness replaced it
    (``monkeypatch.setattr("…commands.images.APIClient", Fake)``) the replacement is built, with no arguments, instead."""
    injected = sys._getframe(1).f_globals.get("APIClient")
    if injected is not None and injected is not APIClient:
        return injected()
    return APIClient(config=Config(writable=False), require_auth=require_auth)


def fail(message: str, code: int = 1) -> "typer.Exit":
    console.print(f"[red]Error:[/red] {message}")
    return typer.Exit(code)


def handle_errors(fn: Callable) -> Callable:
    """Render API failures as one red line (422s field by field) and exit 1 instead of a traceback."""

    @functools.wraps(fn)
    def wrapper(*a: Any, **kw: Any):
        try:
            return fn(*a, **kw)
        except ValidationError as e:
            console.print("[red]Validation error:[/red]")
            for err in e.errors:
                loc = ".".join(str(p) for p in err.get("loc", []) if p != "body")
                console.print(f"  [yellow]{loc}[/yellow]: {err.get('msg', 'invalid')}")
            raise typer.Exit(1)
        except APIError as e:
            raise fail(str(e))
        except KeyboardInterrupt:
            console.print("\n[dim]Cancelled.[/dim]")
            raise typer.Exit(130)
        except typer.Abort:  # a prompt hit EOF or was declined by click itself: click's own convention (and the reference's exit code)
            console.print("\n[dim]Aborted.[/dim]")
            raise typer.Exit(1)
        except (typer.Exit, SystemExit):
            raise
        except Exception as e:  # noqa: BLE001
            # an answer the models do not expect (missing fields, a string where an object should be): one clean line instead of a
            # traceback; PRIME_DEBUG=1 re-raises for the full story
            import os

            import pydantic

            if os.environ.get("PRIME_DEBUG") or not isinstance(e, (pydantic.ValidationError, AttributeError, KeyError, TypeError, IndexError)):
                raise
            what = f"{e.error_count()} field(s) missing or invalid" if isinstance(e, pydantic.ValidationError) else f"{type(e).__name__}: {e}"
            raise fail(f"Unexpected response from the API ({what}). Set PRIME_DEBUG=1 for the traceback.")

    return wrapper


def emit(output: str, payload: Any, title: str | None, columns: Sequence[str | tuple[str, str]],
         rows: Iterable[Sequence[Any]], footer: str | None = None) -> None:  # fmt: skip
    """One call for the list-style commands: JSON document or rich table."""
    if validate_output_format(output, console) == "json":
        output_data_as_json(payload, console)
        return
    console.print(build_table(title, columns, rows))
    if footer:
        console.print(footer)


def paginate_hint(total: int, offset: int, limit: int, noun: str) -> str | None:
    shown_to = min(offset + limit, total)
    if total > shown_to:
        return f"[dim]Showing {offset + 1}-{shown_to} of {total} {noun}. Use --offset {shown_to} for the next page.[/dim]"
    return None
