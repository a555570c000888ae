"""Time helpers: kubectl-style ages, "5m ago", ISO timestamps, sorting by creation time
(reference: packages/prime/src/prime_cli/utils/time_utils.py:9-107)."""

from __future__ import annotations

from datetime import datetime, timezone
from typing import Any, Iterable

ISO_FMT = "%Y-%m-%d %H:%M:%S UTC"
_EPOCH0 = datetime.min.replace(tzinfo=timezone.utc)
_UNITS = ((86400, "d"), (3600, "h"), (60, "m"))


def now_utc() -> datetime:
    return datetime.now(timezone.utc)


def to_utc(dt: datetime) -> datetime:
    return dt if dt.tzinfo else dt.replace(tzinfo=timezone.utc)


def parse_dt(value: datetime | str) -> datetime:
    if isinstance(value, str):
        value = datetime.fromisoformat(value.replace("Z", "+00:00"))
    return to_utc(value)


def _bucket(seconds: int) -> tuple[int, str]:
    for size, unit in _UNITS:
        if seconds >= size:
            return seconds // size, unit
    return seconds, "s"


def human_age(created: datetime | str) -> str:
    n, unit = _bucket(max(0, int((now_utc() - parse_dt(created)).total_seconds())))
    return f"{n}{unit}"


def format_time_ago(dt: datetime | str | None) -> str:
    if not dt:
        return "-"
    when = parse_dt(dt)
    secs = int((now_utc() - when).total_seconds())
    if secs < 60:
        return "just now"
    if secs >= 30 * 86400:
        return when.strftime("%Y-%m-%d")
    n, unit = _bucket(secs)
    return f"{n}{unit} ago"


def iso_timestamp(dt: datetime | str) -> str:
    return parse_dt(dt).strftime(ISO_FMT)


def sort_by_created(items: Iterable[Any], attr: str = "created_at", reverse: bool = False) -> list[Any]:
    def key(item: Any) -> datetime:
        v = getattr(item, attr, None) if not isinstance(item, dict) else item.get(attr)
        try:
            return parse_dt(v) if isinstance(v, (str, datetime)) else _EPOCH0
        except (ValueError, TypeError):
            return _EPOCH0

    return sorted(items, key=key, reverse=reverse)
