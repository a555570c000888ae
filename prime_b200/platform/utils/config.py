"""TOML + CLI merge for job configs: CLI > TOML > defaults, with ``prefix_suffix`` CLI names mapped onto
nested sections (``wandb_project`` → ``[wandb] project``)
(reference: packages/prime/src/prime_cli/utils/config.py:15-111)."""

from __future__ import annotations

import tomllib
from pathlib import Path
from typing import Any

import typer
from pydantic import BaseModel
from typing_extensions import Self

from .plain import get_console


def load_toml(path: str, console=None) -> dict[str, Any]:
    console = console or get_console()
    p = Path(path)
    if not p.exists():
        console.print(f"[red]Error:[/red] Config file not found: {path}")
        raise typer.Exit(1)
    try:
        return tomllib.loads(p.read_text())
    except tomllib.TOMLDecodeError as e:
        console.print(f"[red]Error:[/red] Invalid TOML in {path}: {e}")
        raise typer.Exit(1)


class BaseConfig(BaseModel):
    @classmethod
    def merge_sources(cls, toml_data: dict[str, Any] | None, cli: dict[str, Any]) -> dict[str, Any]:
        data: dict[str, Any] = dict(toml_data or {})
        for key, value in cli.items():
            if value is None:
                continue
            head, _, tail = key.partition("_")
            if key not in cls.model_fields and tail and head in cls.model_fields:
                section = data.setdefault(head, {})
                if isinstance(section, dict):
                    section[tail] = value
                continue
            data[key] = value
        return data

    @classmethod
    def from_sources(cls, toml_path: str | None = None, console=None, **cli_overrides: Any) -> Self:
        toml_data = load_toml(toml_path, console) if toml_path else {}
        return cls.model_validate(cls.merge_sources(toml_data, cli_overrides))
