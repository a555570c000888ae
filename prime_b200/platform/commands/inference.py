"""``prime inference models`` (reference: packages/prime/src/prime_cli/commands/inference.py:51-106)."""

from __future__ import annotations

import datetime as dt

import typer

from ..api.inference import InferenceAPIError, InferenceClient
from ..utils.display import output_data_as_json, validate_output_format
from ..utils.json_help import list_json_help
from ._common import OUTPUT_OPT, console, make_app
from ..utils.display import build_table

app = make_app("Prime Inference (OpenAI-compatible)")


def fmt_created(val) -> str:
    try:
        return dt.datetime.fromtimestamp(int(val), dt.timezone.utc).strftime("%Y-%m-%d %H:%M:%S")
    except (TypeError, ValueError):
        return str(val or "")


def fmt_price(x) -> str:
    """USD per 1M tokens, trailing zeros trimmed."""
    if x is None:
        return ""
    try:
        return f"${float(x):.6f}".rstrip("0").rstrip(".")
    except (TypeError, ValueError):
        return str(x)


def extract_models(data) -> list[dict]:
    if isinstance(data, list):
        return data
    if isinstance(data, dict):
        for key in ("data", "models"):
            if isinstance(data.get(key), list):
                return data[key]
    return []


@app.command("models", epilog=list_json_help("data", {"id": "str", "created": "int", "pricing": {"input_usd_per_mtok": "float", "output_usd_per_mtok": "float"}}))
def list_models(output: str = OUTPUT_OPT) -> None:
    """List models served by Prime Inference."""
    validate_output_format(output, console)
    try:
        data = InferenceClient().list_models()
    except InferenceAPIError as e:
        console.print(f"[red]Error:[/red] {e}")
        raise typer.Exit(1)
    if output == "json":
        return output_data_as_json(data, console)
    models = extract_models(data)
    if not models:
        console.print("[yellow]No models returned.[/yellow]")
        return
    rows = [[m.get("id", ""), fmt_created(m.get("created")), fmt_price((m.get("pricing") or {}).get("input_usd_per_mtok")),
             fmt_price((m.get("pricing") or {}).get("output_usd_per_mtok"))] for m in models]  # fmt: skip
    console.print(build_table("Prime Inference — Models", [("id", "cyan"), ("created", "magenta"), ("input $/1M tok", "green"), ("output $/1M tok", "green")], rows))
