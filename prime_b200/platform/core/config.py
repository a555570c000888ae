"""Identity / endpoint configuration (``~/.prime/config.json``) with env-over-file precedence and
named contexts (``~/.prime/environments/<name>.json``).

Behavioural parity with reference packages/prime/src/prime_cli/core/config.py:10-371:
  * keys: api_key, team_id(+name, role), user_id, base_url, frontend_url, inference_url, ssh_key_path,
    current_environment, share_resources_with_team;
  * env overrides: PRIME_API_KEY, PRIME_TEAM_ID, PRIME_USER_ID, PRIME_API_BASE_URL | PRIME_BASE_URL,
    PRIME_FRONTEND_URL, PRIME_INFERENCE_URL, PRIME_SSH_KEY_PATH; PRIME_CONTEXT selects a saved context
    without persisting it;
  * SDK use is read-only (``Config(writable=False)`` never creates or writes files), which replaces the
    reference's separate read-only copies (prime_sandboxes/core/config.py:9-65 etc.).
"""

from __future__ import annotations

import json
import os
import re
from pathlib import Path
from typing import Any

from pydantic import BaseModel, ConfigDict

DEFAULTS = {
    "base_url": "https://api.primeintellect.ai",
    "frontend_url": "https://app.primeintellect.ai",
    "inference_url": "https://api.pinference.ai/api/v1",
}
BUILTIN_CONTEXT = "production"


def _home() -> Path:
    return Path.home()


class ConfigModel(BaseModel):
    model_config = ConfigDict(populate_by_name=True, extra="ignore")
    api_key: str = ""
    team_id: str | None = None
    team_name: str | None = None
    team_role: str | None = None
    user_id: str | None = None
    base_url: str = DEFAULTS["base_url"]
    frontend_url: str = DEFAULTS["frontend_url"]
    inference_url: str = DEFAULTS["inference_url"]
    ssh_key_path: str = ""
    current_environment: str = BUILTIN_CONTEXT
    share_resources_with_team: bool = False


def normalise_base_url(url: str) -> str:
    """Accept ``https://host/api/v1`` and ``https://host/`` spellings alike."""
    url = url.rstrip("/")
    return url[: -len("/api/v1")] if url.endswith("/api/v1") else url


def sanitise_context_name(name: str) -> str:
    cleaned = re.sub(r"[^A-Za-z0-9_-]", "_", name.strip())
    if not cleaned:
        raise ValueError("context name must contain at least one letter or digit")
    return cleaned


class Config:
    ENV_KEYS = {
        "api_key": ("PRIME_API_KEY",),
        "team_id": ("PRIME_TEAM_ID",),
        "user_id": ("PRIME_USER_ID",),
        "base_url": ("PRIME_API_BASE_URL", "PRIME_BASE_URL"),
        "frontend_url": ("PRIME_FRONTEND_URL",),
        "inference_url": ("PRIME_INFERENCE_URL",),
        "ssh_key_path": ("PRIME_SSH_KEY_PATH",),
    }

    def __init__(self, writable: bool = True) -> None:
        self.writable = writable
        self.config_dir = _home() / ".prime"
        self.config_file = self.config_dir / "config.json"
        self.environments_dir = self.config_dir / "environments"
        self.data: dict[str, Any] = ConfigModel(ssh_key_path=self.default_ssh_key_path()).model_dump()
        if writable:
            self.config_dir.mkdir(parents=True, exist_ok=True)
            self.environments_dir.mkdir(exist_ok=True)
            if not self.config_file.exists():
                self._write(self.data)
        self._read()
        ctx = os.getenv("PRIME_CONTEXT")
        if ctx:
            self.load_environment(ctx, persist=False)

    # ------------------------------------------------------------------ file io
    @staticmethod
    def default_ssh_key_path() -> str:
        return str(_home() / ".ssh" / "id_rsa")

    def _read(self) -> None:
        if self.config_file.exists():
            try:
                raw = json.loads(self.config_file.read_text() or "{}")
            except json.JSONDecodeError:
                raw = {}
            merged = ConfigModel(**{**{"ssh_key_path": self.default_ssh_key_path()}, **raw})
            self.data = merged.model_dump()

    def _write(self, data: dict[str, Any]) -> None:
        if not self.writable:
            raise PermissionError("this Config instance is read-only (SDK mode)")
        tmp = self.config_file.with_suffix(".tmp")
        tmp.write_text(json.dumps(data, indent=2))
        os.replace(tmp, self.config_file)
        self.data = data

    def _set(self, key: str, value: Any) -> None:
        self.data[key] = value
        self._write(self.data)

    def _env(self, key: str) -> str | None:
        for name in self.ENV_KEYS.get(key, ()):
            v = os.getenv(name)
            if v is not None and v.strip():
                return v
        return None

    # ------------------------------------------------------------------ properties (env > file > default)
    @property
    def api_key(self) -> str:
        return self._env("api_key") or self.data.get("api_key", "") or ""

    def set_api_key(self, value: str) -> None:
        self._set("api_key", value)

    @property
    def team_id(self) -> str | None:
        return self._env("team_id") or (self.data.get("team_id") or None)

    @property
    def team_id_from_env(self) -> bool:
        return self._env("team_id") is not None

    @property
    def team_name(self) -> str | None:
        return self.data.get("team_name") or None

    @property
    def team_role(self) -> str | None:
        return self.data.get("team_role") or None

    def set_team(self, value: str | None = None, team_name: str | None = None, team_role: str | None = None, *, team_id: str | None = None) -> None:
        team_id = value if value is not None else team_id  # `value` is the reference's parameter name; `team_id` kept for callers
        self.data["team_id"] = team_id or None
        self.data["team_name"] = team_name if team_id else None
        self.data["team_role"] = team_role if team_id else None
        self._write(self.data)

    def set_team_id(self, value: str | None) -> None:
        self.set_team(value)

    @property
    def user_id(self) -> str | None:
        return self._env("user_id") or (self.data.get("user_id") or None)

    def set_user_id(self, value: str | None) -> None:
        self._set("user_id", value or None)

    @property
    def base_url(self) -> str:
        return normalise_base_url(self._env("base_url") or self.data.get("base_url") or DEFAULTS["base_url"])

    def set_base_url(self, value: str) -> None:
        self._set("base_url", normalise_base_url(value))

    @property
    def frontend_url(self) -> str:
        return (self._env("frontend_url") or self.data.get("frontend_url") or DEFAULTS["frontend_url"]).rstrip("/")

    def set_frontend_url(self, value: str) -> None:
        self._set("frontend_url", value.rstrip("/"))

    @property
    def inference_url(self) -> str:
        return (self._env("inference_url") or self.data.get("inference_url") or DEFAULTS["inference_url"]).rstrip("/")

    def set_inference_url(self, value: str) -> None:
        self._set("inference_url", value.rstrip("/"))

    @property
    def ssh_key_path(self) -> str:
        return self._env("ssh_key_path") or self.data.get("ssh_key_path") or self.default_ssh_key_path()

    def set_ssh_key_path(self, value: str) -> None:
        self._set("ssh_key_path", str(Path(value).expanduser()))

    @property
    def current_environment(self) -> str:
        return self.data.get("current_environment") or BUILTIN_CONTEXT

    def set_current_environment(self, value: str) -> None:
        """Remember which named context is active (reference: packages/prime/src/prime_cli/core/config.py:212-215)."""
        self._set("current_environment", value)

    @property
    def share_resources_with_team(self) -> bool:
        return bool(self.data.get("share_resources_with_team", False))

    def set_share_resources_with_team(self, value: bool) -> None:
        self._set("share_resources_with_team", bool(value))

    @property
    def bin_dir(self) -> Path:
        return self.config_dir / "bin"

    def view(self) -> dict[str, Any]:
        return {
            "api_key": self.api_key,
            "team_id": self.team_id,
            "team_name": self.team_name,
            "user_id": self.user_id,
            "base_url": self.base_url,
            "frontend_url": self.frontend_url,
            "inference_url": self.inference_url,
            "ssh_key_path": self.ssh_key_path,
            "current_environment": self.current_environment,
            "share_resources_with_team": self.share_resources_with_team,
        }

    def reset(self) -> None:
        """``prime config reset``: credentials, team, API / frontend URLs, SSH key path and the active context go back to their defaults;
        who you are (``user_id``), the inference URL and the sharing preference stay (reference: commands/config.py:363-375)."""
        self.data.update(api_key="", team_id=None, team_name=None, team_role=None, base_url=DEFAULTS["base_url"],
                         frontend_url=DEFAULTS["frontend_url"], ssh_key_path=self.default_ssh_key_path(), current_environment=BUILTIN_CONTEXT)  # fmt: skip
        self._write(self.data)

    # ------------------------------------------------------------------ named contexts
    # the reference's class-level names for the defaults (read by its tunnel / CLI tests and by user code)
    DEFAULT_BASE_URL = DEFAULTS["base_url"]
    DEFAULT_FRONTEND_URL = DEFAULTS["frontend_url"]
    DEFAULT_INFERENCE_URL = DEFAULTS["inference_url"]

    CONTEXT_FIELDS = ("api_key", "team_id", "team_name", "team_role", "user_id", "base_url", "frontend_url",
                      "inference_url", "share_resources_with_team")  # fmt: skip

    def _context_path(self, name: str) -> Path:
        return self.environments_dir / f"{sanitise_context_name(name)}.json"

    def save_environment(self, name: str) -> str:
        clean = sanitise_context_name(name)
        if clean == BUILTIN_CONTEXT:
            raise ValueError(f"'{BUILTIN_CONTEXT}' is built in and cannot be overwritten")
        self.environments_dir.mkdir(parents=True, exist_ok=True)
        self._context_path(clean).write_text(json.dumps(self._context_snapshot(), indent=2))
        return clean

    def _context_snapshot(self) -> dict[str, Any]:
        """What a named context stores: the EFFECTIVE settings — an exported ``PRIME_API_KEY`` / ``PRIME_API_BASE_URL`` is what the user is
        working with and is what gets saved (reference: core/config.py:244-261); team name / role only when the team is not env-given."""
        team_from_env = self._env("team_id") is not None
        return {"api_key": self.api_key, "team_id": self.team_id, "team_name": None if team_from_env else self.team_name,
                "team_role": None if team_from_env else self.team_role, "user_id": self.user_id, "base_url": self.base_url,
                "frontend_url": self.frontend_url, "inference_url": self.inference_url,
                "share_resources_with_team": self.share_resources_with_team}  # fmt: skip

    def update_current_environment_file(self) -> None:
        """Keep the active named context in sync with edits made while it is selected."""
        cur = self.current_environment
        if cur != BUILTIN_CONTEXT and self._context_path(cur).exists():
            self._context_path(cur).write_text(json.dumps(self._context_snapshot(), indent=2))

    def list_environments(self) -> list[str]:
        names = {BUILTIN_CONTEXT}
        if self.environments_dir.exists():
            names.update(p.stem for p in self.environments_dir.glob("*.json"))
        return sorted(names)

    def load_environment(self, name: str, persist: bool = True) -> bool:
        clean = sanitise_context_name(name)
        if clean == BUILTIN_CONTEXT:
            patch = {**DEFAULTS, "team_id": None, "team_name": None, "team_role": None}
            if not persist:
                # temporary switch keeps the credentials already loaded
                patch = dict(DEFAULTS)
        else:
            path = self._context_path(clean)
            if not path.exists():
                return False
            try:
                patch = json.loads(path.read_text())
            except json.JSONDecodeError:
                return False
        new = {**self.data, **{k: v for k, v in patch.items() if k in ConfigModel.model_fields}}
        new["current_environment"] = clean
        if persist:
            self._write(new)
        else:
            self.data = new
        return True

    def delete_environment(self, name: str) -> bool:
        clean = sanitise_context_name(name)
        if clean == BUILTIN_CONTEXT:
            raise ValueError(f"'{BUILTIN_CONTEXT}' cannot be deleted")
        path = self._context_path(clean)
        if not path.exists():
            return False
        path.unlink()
        return True
