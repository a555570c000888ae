"""Hosted evaluations, client side: run states, the request a run is created from, and turning the raw container log
(ANSI colours, tqdm redraws, placeholder messages while the container boots) into lines worth printing — including
"what is new since the last poll" for ``--follow``.

Behaviour matches the reference's helpers (packages/prime/src/prime_cli/utils/hosted_eval.py:12-112); the structure is
ours: one classifier for tqdm lines (shared with ``rl logs``), the wire payload built by the config itself, and the
poll-to-poll overlap found longest-first.
"""

from __future__ import annotations

import re
from dataclasses import dataclass, fields
from enum import Enum
from typing import Any

from .formatters import strip_ansi

# ---------------------------------------------------------------------------------------------------------- run states
_TERMINAL = frozenset({"COMPLETED", "FAILED", "TIMEOUT", "CANCELLED"})
_COLOURS = {"PENDING": "yellow", "RUNNING": "cyan", "COMPLETED": "green", "CANCELLED": "yellow", "FAILED": "red", "TIMEOUT": "red"}


class EvalStatus(str, Enum):
    PENDING = "PENDING"
    RUNNING = "RUNNING"
    COMPLETED = "COMPLETED"
    FAILED = "FAILED"
    TIMEOUT = "TIMEOUT"
    CANCELLED = "CANCELLED"

    @property
    def is_terminal(self) -> bool:
        return self.value in _TERMINAL

    @classmethod
    def terminal_statuses(cls) -> set["EvalStatus"]:
        return {s for s in cls if s.is_terminal}

    @property
    def color(self) -> str:
        return _COLOURS.get(self.value, "white")


# ------------------------------------------------------------------------------------------------------- run request
@dataclass
class HostedEvalConfig:
    """Everything ``prime eval run --hosted`` sends. ``payload()`` is the POST body."""

    environment_id: str
    inference_model: str
    num_examples: int
    rollouts_per_example: int
    env_args: dict[str, str] | None = None
    name: str | None = None
    timeout_minutes: int | None = None
    allow_sandbox_access: bool = False
    allow_instances_access: bool = False
    custom_secrets: dict[str, str] | None = None
    sampling_args: dict[str, Any] | None = None
    api_base_url: str | None = None
    api_key_var: str | None = None

    _TOP_LEVEL = ("environment_id", "inference_model", "name")
    _SEND_IF_TRUTHY = ("env_args", "custom_secrets", "sampling_args", "api_base_url", "api_key_var")

    def payload(self) -> dict[str, Any]:
        run: dict[str, Any] = {}
        for f in fields(self):
            value = getattr(self, f.name)
            if f.name in self._TOP_LEVEL:
                continue
            if f.name in self._SEND_IF_TRUTHY and not value:
                continue
            if value is None:
                continue
            run[f.name] = value
        body: dict[str, Any] = {"environment_ids": [self.environment_id], "inference_model": self.inference_model, "eval_config": run}
        if self.name:
            body["name"] = self.name
        return body


@dataclass
class HostedEvalResult:
    evaluation_id: str
    status: EvalStatus
    total_samples: int
    avg_score: float | None
    min_score: float | None
    max_score: float | None
    error_message: str | None = None
    logs: str | None = None


# ----------------------------------------------------------------------------------------------------------- log text
_BLOCKS = "█▏▎▍▌▋▊▉ "
PROGRESS_BAR = re.compile(rf".*\|[{_BLOCKS}]{{10,}}\|.*")  # a drawn bar at least ten cells wide
_TQDM_HEAD = re.compile(r"\d+%\|")  # "NN%|" — tqdm's prefix, whatever the bar is drawn with
_FINISHED = re.compile(rf"([^|]*100%\|[{_BLOCKS}]+\|[^\n]*?)(?=\d+%\||$)")

STATUS_MESSAGES = ("Waiting for container to start...", "No logs available", "Unable to retrieve logs",
                   "Failed to fetch logs from sandbox", "The hosted eval is still initializing")  # fmt: skip


def tqdm_line(line: str) -> str | None:
    """Classify one ANSI-free line: ``None`` = not tqdm output; ``""`` = an intermediate redraw (drop it);
    otherwise the text of the finished bar (several redraws can share a line when ``\\r`` was flattened)."""
    if not (PROGRESS_BAR.search(line) or _TQDM_HEAD.search(line)):
        return None
    if "100%" not in line:
        return ""
    done = _FINISHED.search(line)
    return (done.group(1) if done else line).strip()


def filter_progress_bars(text: str) -> str:
    kept = []
    for line in text.splitlines():
        bar = tqdm_line(line)
        if bar is None:
            if line.strip():
                kept.append(line)
        elif bar:
            kept.append(bar)
    return "\n".join(kept)


def is_status_message(text: str) -> bool:
    return text.strip().startswith(STATUS_MESSAGES)


def clean_logs(text: str) -> str:
    body = filter_progress_bars(strip_ansi(text))
    return "" if is_status_message(body) else body


def get_new_log_lines(old_logs: str, new_logs: str) -> list[str]:
    """The API returns a sliding tail window, so consecutive polls overlap: skip the longest prefix of the new window
    that is a suffix of the old one."""
    fresh = new_logs.splitlines()
    seen = old_logs.splitlines() if old_logs else []
    for n in range(min(len(seen), len(fresh)), 0, -1):
        if seen[-n:] == fresh[:n]:
            return fresh[n:]
    return fresh
