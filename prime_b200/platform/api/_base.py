"""Shared model base and error wrapping for the typed domain clients."""

from __future__ import annotations

from contextlib import contextmanager
from typing import Any, Iterable

from pydantic import BaseModel, ConfigDict
from pydantic.alias_generators import to_camel

from ..core.client import APIError


class ApiModel(BaseModel):
    """snake_case in Python, camelCase on the wire; unknown wire fields are kept (forward compatible)."""

    model_config = ConfigDict(alias_generator=to_camel, populate_by_name=True, extra="allow")

    def __getitem__(self, key: str) -> Any:
        return getattr(self, key)


@contextmanager
def wrap(action: str):
    """Re-raise anything as ``APIError("Failed to <action>: …")`` keeping typed subclasses intact."""
    try:
        yield
    except APIError as e:
        if type(e) is not APIError:
            raise
        raise APIError(f"Failed to {action}: {e}", e.status_code) from e
    except Exception as e:  # pydantic validation, json, …
        raise APIError(f"Failed to {action}: {e}") from e


def split_multi(values: Iterable[str] | None) -> list[str]:
    """``["a,b", "c"]`` → ``["a", "b", "c"]`` (repeatable, comma-separable CLI options)."""
    out: list[str] = []
    for v in values or ():
        out.extend(p.strip() for p in v.split(",") if p.strip())
    return out


def unwrap_single_none(v: Any) -> Any:
    """API quirk: ``[None]`` means "no value"."""
    if isinstance(v, list) and len(v) == 1 and v[0] is None:
        return None
    return v
