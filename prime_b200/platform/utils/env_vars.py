"""Environment-variable / secret collection for job submission.

Same contract as reference packages/prime/src/prime_cli/utils/env_vars.py:39-182:
``.env`` files (``KEY=VALUE``, quotes, ``#`` comments, ``${VAR}`` expansion unless single-quoted),
``-e KEY=VALUE`` / ``-e KEY`` (value taken from the process env) / ``-e path.env``; files are applied
first, explicit ``-e`` arguments override them.  Implemented as a tiny line tokenizer + one merge.
"""

from __future__ import annotations

import os
import re
from pathlib import Path
from typing import Callable, Iterable

Warn = Callable[[str], None] | None

_KEY = re.compile(r"[A-Za-z_][A-Za-z0-9_]*\Z")
_REF = re.compile(r"\$\{([A-Za-z_][A-Za-z0-9_]*)\}")
_KEY_RULE = "must start with a letter or underscore and contain only letters, digits and underscores"


class EnvParseError(Exception):
    pass


def _unquote(raw: str) -> tuple[str, bool]:
    """Returns (value, literal) — ``literal`` is True for single-quoted values (no expansion)."""
    if len(raw) >= 2 and raw[0] == raw[-1] and raw[0] in "'\"":
        return raw[1:-1], raw[0] == "'"
    return raw, False


def _expand(value: str, where: str) -> str:
    def sub(m: re.Match[str]) -> str:
        got = os.environ.get(m.group(1))
        if got is None:
            raise EnvParseError(f"Environment variable '{m.group(1)}' is not set (referenced in {where}).")
        return got

    return _REF.sub(sub, value)


def parse_env_file(file_path: Path, on_warning: Warn = None) -> dict[str, str]:
    out: dict[str, str] = {}
    warn = on_warning or (lambda _m: None)
    for n, raw in enumerate(Path(file_path).read_text().splitlines(), 1):
        line = raw.strip()
        if not line or line.startswith("#"):
            continue
        key, sep, rest = line.partition("=")
        if not sep:
            warn(f"Skipping invalid line {n} in {file_path}: missing '=' separator")
            continue
        key = key.strip()
        if not _KEY.match(key):
            warn(f"Skipping invalid key '{key}' in {file_path}: {_KEY_RULE}")
            continue
        value, literal = _unquote(rest.strip())
        out[key] = value if literal else _expand(value, f"{file_path}:{n}")
    return out


def parse_env_arg(arg: str, on_warning: Warn = None) -> dict[str, str]:
    if Path(arg).is_file():
        return parse_env_file(Path(arg), on_warning)
    key, sep, rest = arg.partition("=")
    key = key.strip()
    if not _KEY.match(key):
        raise EnvParseError(f"Invalid environment variable key '{key}': {_KEY_RULE}")
    if sep:
        return {key: _unquote(rest.strip())[0]}
    if key not in os.environ:
        raise EnvParseError(f"Environment variable '{key}' is not set. Either set it or use KEY=VALUE syntax.")
    return {key: os.environ[key]}


def collect_env_vars(env_args: Iterable[str] | None = None, env_files: Iterable[str] | None = None,
                     on_warning: Warn = None) -> dict[str, str]:  # fmt: skip
    merged: dict[str, str] = {}
    for f in env_files or ():
        if not Path(f).is_file():
            raise EnvParseError(f"Env file not found: {f}")
        merged.update(parse_env_file(Path(f), on_warning))
    for a in env_args or ():
        merged.update(parse_env_arg(a, on_warning))
    return merged
