"""Interactive prompt helpers (reference: packages/prime/src/prime_cli/utils/prompt.py:13-127)."""

from __future__ import annotations

import re
from typing import Any, Callable

import typer

from .plain import get_console

console = get_console()
_SECRET_NAME = re.compile(r"[A-Z][A-Z0-9_]*\Z")


def validate_env_var_name(name: str, item_type: str = "secret") -> bool:
    if _SECRET_NAME.match(name):
        return True
    console.print(f"[red]Invalid {item_type} name: '{name}'[/red]")
    console.print(f"[dim]{item_type.capitalize()} names are UPPER_SNAKE_CASE: uppercase letters, digits and "
                  "underscores, starting with a letter (e.g. MY_SECRET, API_KEY_2).[/dim]")  # fmt: skip
    return False


def confirm_or_skip(message: str, yes_flag: bool, default: bool = False) -> bool:
    return True if yes_flag else bool(typer.confirm(message, default=default))


def any_provided(*values: Any) -> bool:
    return any(v is not None for v in values)


def _label(item: dict[str, Any]) -> str:
    name, desc = item.get("name", ""), item.get("description") or ""
    return f"{name} - {desc}" if desc else str(name)


def select_item_interactive(items: list[dict[str, Any]], action: str = "select", item_type: str = "item",
                            display_fn: Callable[[dict[str, Any]], str] | None = None) -> dict[str, Any] | None:  # fmt: skip
    if not items:
        return None
    console.print(f"\n[bold]Select a {item_type} to {action}:[/bold]\n")
    for i, it in enumerate(items, 1):
        console.print(f"  {i}. {(display_fn or _label)(it)}")
    console.print()
    while True:
        try:
            raw = typer.prompt("Select (empty to cancel)", default="", show_default=False)
        except (KeyboardInterrupt, typer.Abort):
            return None
        if not raw:
            return None
        if raw.isdigit() and 1 <= int(raw) <= len(items):
            return items[int(raw) - 1]
        console.print(f"[red]Please enter a number between 1 and {len(items)}[/red]")


def require_selection(items: list[dict[str, Any]], action: str, empty_message: str, item_type: str = "item",
                      display_fn: Callable[[dict[str, Any]], str] | None = None) -> dict[str, Any]:  # fmt: skip
    if not items:
        console.print(f"[yellow]{empty_message}[/yellow]")
        raise typer.Exit()
    chosen = select_item_interactive(items, action, item_type, display_fn)
    if chosen is None:
        console.print("\n[dim]Cancelled.[/dim]")
        raise typer.Exit()
    return chosen


def prompt_for_value(prompt_text: str, required: bool = True, hide_input: bool = False) -> str | None:
    try:
        v = typer.prompt(prompt_text + (" (empty to cancel)" if required else ""), default="", hide_input=hide_input,
                         show_default=False)  # fmt: skip
    except KeyboardInterrupt:
        return None
    # an EMPTY answer cancels (exit 0); a closed stdin is click's Abort and stays one ("Aborted.", exit 1 — scripts that forgot a flag
    # must not see success)
    return None if (required and not v) else v
