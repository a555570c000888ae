"""Locate an environment's ``.prime/.env-metadata.json`` (legacy: ``.env-metadata.json`` at the root)
(reference: packages/prime/src/prime_cli/utils/env_metadata.py:8-86)."""

from __future__ import annotations

import json
from pathlib import Path
from typing import Any

NEW_REL = Path(".prime") / ".env-metadata.json"
LEGACY_REL = Path(".env-metadata.json")


def get_environment_metadata(env_path: Path) -> dict[str, Any] | None:
    for rel in (NEW_REL, LEGACY_REL):
        f = Path(env_path) / rel
        if f.exists():
            try:
                return json.loads(f.read_text())
            except (json.JSONDecodeError, OSError):
                return None
    return None


def candidate_dirs(env_name: str | None = None, env_path: Path | None = None, module_name: str | None = None) -> list[Path]:
    dirs: list[Path] = []
    if env_path:
        dirs.append(Path(env_path))
    if module_name:
        dirs.append(Path("environments") / module_name)
    if env_name:
        dirs += [Path("environments") / env_name, Path(env_name)]
    if module_name:
        dirs.append(Path(module_name))
    dirs.append(Path("."))
    return dirs


def find_environment_metadata(env_name: str | None = None, env_path: Path | None = None,
                              module_name: str | None = None) -> dict[str, Any] | None:  # fmt: skip
    for d in candidate_dirs(env_name, env_path, module_name):
        md = get_environment_metadata(d)
        if md:
            return md
    return None


def write_environment_metadata(env_path: Path, metadata: dict[str, Any]) -> Path:
    """Write to the new location and remove a legacy root-level file if present (migration)."""
    target = Path(env_path) / NEW_REL
    target.parent.mkdir(parents=True, exist_ok=True)
    target.write_text(json.dumps(metadata, indent=2))
    legacy = Path(env_path) / LEGACY_REL
    if legacy.exists():
        legacy.unlink()
    return target
