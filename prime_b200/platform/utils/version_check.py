"""Non-blocking "a newer release exists" banner: PyPI poll with a 24 h cache, at most one notification per
6 h, 2 s network budget, disabled by ``PRIME_DISABLE_VERSION_CHECK``
(reference: packages/prime/src/prime_cli/utils/version_check.py:12-137)."""

from __future__ import annotations

import json
import os
import time
from pathlib import Path

import httpx
from packaging.version import InvalidVersion, Version

PYPI_URL = "https://pypi.org/pypi/prime/json"
CACHE_TTL_S = 24 * 3600
RENOTIFY_S = 6 * 3600
TIMEOUT_S = 2.0


def _cache_path() -> Path:
    d = Path.home() / ".prime"
    d.mkdir(exist_ok=True)
    return d / "version_check.json"


def _load() -> dict:
    try:
        return json.loads(_cache_path().read_text())
    except (OSError, json.JSONDecodeError):
        return {}


def _store(state: dict) -> None:
    try:
        _cache_path().write_text(json.dumps(state))
    except OSError:
        pass


def get_latest_pypi_version() -> str | None:
    try:
        r = httpx.get(PYPI_URL, timeout=TIMEOUT_S)
        r.raise_for_status()
        return r.json()["info"]["version"]
    except Exception:
        return None


def check_for_update(installed: str | None = None, now: float | None = None) -> tuple[bool, str | None]:
    """Returns (should_notify_now, latest_version_or_None)."""
    if os.environ.get("PRIME_DISABLE_VERSION_CHECK", "").lower() in ("1", "true", "yes"):
        return False, None
    from .. import __version__

    now = time.time() if now is None else now
    try:
        state = _load()
        fresh = now - state.get("last_check", 0) < CACHE_TTL_S and state.get("latest_version")
        if not fresh:
            latest = get_latest_pypi_version()
            if not latest:
                return False, None
            state.update(last_check=now, latest_version=latest)
            _store(state)
        latest = state["latest_version"]
        if Version(installed or __version__) >= Version(latest):
            return False, latest
        last = state.get("last_notified")
        if last is not None and now - last < RENOTIFY_S:
            return False, latest
        state["last_notified"] = now
        _store(state)
        return True, latest
    except (InvalidVersion, Exception):
        return False, None
