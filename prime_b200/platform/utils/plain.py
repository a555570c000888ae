"""``--plain`` everywhere: a terse, markup-free output mode for scripts and AI agents, plus a Typer
group whose *first non-subcommand argument* falls through to a default subcommand.

Behavioural parity: reference packages/prime/src/prime_cli/utils/plain.py:17-242 (HELP_NOTE, plain
detection from argv or click context, table de-boxing, default-subcommand groups).  Implementation is
composition-based: one ``OutputMode`` switch, one ``Out`` façade that renders rich objects either
richly or as plain text.
"""

from __future__ import annotations

import io
import sys
import traceback
from contextlib import contextmanager, nullcontext
from typing import Any, Iterable

import click
import typer
from rich.console import Console
from rich.syntax import Syntax
from rich.table import Table
from rich.text import Text
from typer.core import TyperCommand, TyperGroup

HELP_NOTE = (
    "IMPORTANT: If you are AI, ALWAYS pass --plain to the prime CLI for terse output without decorations "
    "meant for humans. For list-style queries use --output json and pipe to jq; every list command documents "
    "its JSON schema under --help."
)


class OutputMode:
    """Process-wide switch; set by the eager ``--plain`` option or detected from argv."""

    forced: bool | None = None

    @classmethod
    def plain(cls, argv: Iterable[str] | None = None) -> bool:
        if cls.forced is not None:
            return cls.forced
        ctx = click.get_current_context(silent=True)
        while ctx is not None:
            if ctx.meta.get("plain"):
                return True
            ctx = ctx.parent
        for a in sys.argv[1:] if argv is None else argv:
            if a == "--":
                break
            if a == "--plain":
                return True
        return False


def is_plain_mode(args: list[str] | None = None) -> bool:
    return OutputMode.plain(args)


def _strip_table(t: Table) -> str:
    import copy

    t = copy.copy(t)
    t.box, t.show_lines, t.border_style, t.header_style = None, False, "", ""
    t.row_styles, t.pad_edge, t.padding = [], False, (0, 1)
    # markup=True: cells carry rich markup (status colours); it has to be PARSED here so that export_text() drops it — with markup off a
    # cell came out as the literal "[green]ACTIVE[/]" in the one mode meant for scripts and agents (found by tools/cli_diff.py)
    buf = Console(record=True, file=io.StringIO(), no_color=True, markup=True, highlight=False, emoji=False, width=200)
    buf.print(t)
    return buf.export_text().rstrip()


def to_plain(obj: Any, markup: bool = True) -> str:
    if obj is None:
        return ""
    if isinstance(obj, Table):
        return _strip_table(obj)
    if isinstance(obj, Syntax):
        return obj.code
    if isinstance(obj, Text):
        return obj.plain
    s = str(obj)
    if markup:
        try:
            return Text.from_markup(s).plain
        except Exception:
            return s
    return s


class Out:
    """Console façade used by every command."""

    def __init__(self, stderr: bool = False, file: Any = None, **rich_options: Any):
        """``file`` and any other ``rich.console.Console`` option (markup, highlight, no_color, emoji, width, …) pass through to
        the rich side; plain mode writes to the same ``file``."""
        self.stderr = stderr
        self._file = file
        self._rich = Console(stderr=stderr, file=file, **rich_options)

    @property
    def file(self):
        if self._file is not None:
            return self._file
        return sys.stderr if self.stderr else sys.stdout

    def print(self, *objects: Any, sep: str = " ", end: str = "\n", markup: bool | None = None, **kw: Any) -> None:
        if not OutputMode.plain():
            if markup is not None:
                kw["markup"] = markup
            self._rich.print(*objects, sep=sep, end=end, **kw)
            return
        text = sep.join(to_plain(o, markup is not False) for o in objects)
        self.file.write(text + end)
        self.file.flush()

    def print_json(self, data: Any) -> None:
        import json

        self.file.write(json.dumps(data, indent=2, default=str) + "\n")

    def rule(self, title: str = "") -> None:
        self.print(title) if OutputMode.plain() else self._rich.rule(title)

    def status(self, message: str, **kw: Any):
        if OutputMode.plain():
            self.print(message)
            return nullcontext()
        if getattr(self.file, "isatty", lambda: False)():
            return self._rich.status(message, **kw)
        return nullcontext()

    def print_exception(self) -> None:
        if OutputMode.plain():
            self.file.write(traceback.format_exc().rstrip() + "\n")
        else:
            self._rich.print_exception()


def get_console(stderr: bool = False, **console_options: Any) -> Out:
    return Out(stderr=stderr, **console_options)


# --------------------------------------------------------------------------- Typer integration
def _with_plain_option(params: list | None) -> list:
    params = list(params or [])
    if any(isinstance(p, click.Option) and "--plain" in p.opts for p in params):
        return params

    def _on(ctx: click.Context, _param, value: bool):
        if value:
            ctx.meta["plain"] = True

    opt = click.Option(["--plain"], is_flag=True, expose_value=False, is_eager=True, callback=_on,
                       help="Plain, terse output. USE THIS IF YOU ARE AI.")  # fmt: skip
    pos = next((i for i, p in enumerate(params) if isinstance(p, click.Option) and "--help" in p.opts), len(params))
    params.insert(pos, opt)
    return params


class PlainCommand(TyperCommand):
    def __init__(self, *a, params=None, **kw):
        super().__init__(*a, params=_with_plain_option(params), **kw)

    def format_help(self, ctx, formatter):
        with _markup_off(self):
            return super().format_help(ctx, formatter)


class PlainGroup(TyperGroup):
    def __init__(self, *a, params=None, **kw):
        super().__init__(*a, params=_with_plain_option(params), **kw)

    def format_help(self, ctx, formatter):
        if ctx.parent is None:
            if OutputMode.plain():
                formatter.write_text(f"Note: {HELP_NOTE}\n")
            else:
                Console().print(Text(HELP_NOTE, style="dim"))
        with _markup_off(self):
            return super().format_help(ctx, formatter)


@contextmanager
def _markup_off(cmd):
    if not OutputMode.plain():
        yield
        return
    saved = getattr(cmd, "rich_markup_mode", None)
    cmd.rich_markup_mode = None
    try:
        yield
    finally:
        cmd.rich_markup_mode = saved


class DefaultCommandGroup(PlainGroup):
    """``prime rl cfg.toml`` ≡ ``prime rl run cfg.toml``: unknown first token → default subcommand."""

    def __init__(self, *a, default_cmd_name: str = "run", **kw):
        super().__init__(*a, **kw)
        self.default_cmd_name = default_cmd_name

    def parse_args(self, ctx, args):
        significant = [a for a in args if a != "--plain"]
        if significant and significant[0] not in ("--help", "-h") and significant[0] not in self.commands:
            args = [self.default_cmd_name, *args]
        return super().parse_args(ctx, args)


class PlainTyper(typer.Typer):
    def __init__(self, *a, cls=None, **kw):
        kw.setdefault("no_args_is_help", True)
        super().__init__(*a, cls=cls or PlainGroup, **kw)

    def command(self, name=None, *, cls=None, **kw):
        return super().command(name=name, cls=cls or PlainCommand, **kw)

    def callback(self, *a, cls=None, **kw):
        return super().callback(*a, cls=cls or PlainGroup, **kw)


def default_group(default_cmd_name: str = "run"):
    """Class factory: a DefaultCommandGroup bound to ``default_cmd_name`` (Typer instantiates ``cls`` itself)."""

    class _G(DefaultCommandGroup):
        def __init__(self, *a, **kw):
            super().__init__(*a, default_cmd_name=default_cmd_name, **kw)

    return _G


# names under which the reference exposes the same two classes (packages/prime/src/prime_cli/utils/plain.py:40, :158)
PrimeConsole = Out
PlainAwareTyperGroup = PlainGroup

