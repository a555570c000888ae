"""Epilog builders documenting ``--output json`` shapes
(reference: packages/prime/src/prime_cli/utils/json_help.py:6-18)."""

from __future__ import annotations

import json
from typing import Any


def json_output_help(shape: Any, note: str | None = None) -> str:
    lines = ["JSON output (--output json):", "", json.dumps(shape, indent=2)]
    if note:
        lines += ["", note]
    return "\n".join(lines)


def list_json_help(key: str, item_fields: dict[str, str], extra: dict[str, str] | None = None) -> str:
    shape: dict[str, Any] = {key: [item_fields]}
    shape.update(extra or {})
    return json_output_help(shape)


def json_help(*lines: str) -> str:
    """Epilog for commands that ALWAYS print JSON (reference: packages/prime/src/prime_cli/utils/json_help.py:16-18)."""
    return "\n".join(["JSON output:", "", *lines])

