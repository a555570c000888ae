"""Hosted-evaluation helpers: status enum, request config, log cleaning (ANSI + tqdm bars), incremental log diffing
(reference: packages/prime/src/prime_cli/utils/hosted_eval.py:12-112)."""

from __future__ import annotations

import re
from dataclasses import dataclass
from enum import Enum
from typing import Any

from .formatters import strip_ansi

_BAR_CHARS = "█▏▎▍▌▋▊▉ "
PROGRESS_BAR = re.compile(rf".*\|[{_BAR_CHARS}]{{10,}}\|.*")
_PCT = re.compile(r"\d+%\|")
_DONE = re.compile(rf"([^|]*100%\|[{_BAR_CHARS}]+\|[^\n]*?)(?=\d+%\||$)")
STATUS_MESSAGES = ("Waiting for container to start...", "No logs available", "Unable to retrieve logs",
                   "Failed to fetch logs from sandbox", "The hosted eval is still initializing")  # fmt: skip


class EvalStatus(str, Enum):
    PENDING = "PENDING"
    RUNNING = "RUNNING"
    COMPLETED = "COMPLETED"
    FAILED = "FAILED"
    TIMEOUT = "TIMEOUT"
    CANCELLED = "CANCELLED"

    @classmethod
    def terminal_statuses(cls) -> set["EvalStatus"]:
        return {cls.COMPLETED, cls.FAILED, cls.TIMEOUT, cls.CANCELLED}

    @property
    def color(self) -> str:
        return {"PENDING": "yellow", "RUNNING": "cyan", "COMPLETED": "green", "CANCELLED": "yellow"}.get(self.value, "red")


@dataclass
class HostedEvalConfig:
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


def filter_progress_bars(text: str) -> str:
    """Drop tqdm refresh lines; keep only the final 100 % rendering of each bar."""
    kept: list[str] = []
    for line in text.splitlines():
        if PROGRESS_BAR.search(line) or _PCT.search(line):
            if "100%" in line:
                m = _DONE.search(line)
                kept.append((m.group(1) if m else line).strip())
        elif line.strip():
            kept.append(line)
    return "\n".join(kept)


def is_status_message(text: str) -> bool:
    return text.strip().startswith(STATUS_MESSAGES)


def clean_logs(text: str) -> str:
    cleaned = filter_progress_bars(strip_ansi(text))
    return "" if is_status_message(cleaned) else cleaned


def get_new_log_lines(old_logs: str, new_logs: str) -> list[str]:
    """Lines of ``new_logs`` not already shown: longest suffix of old == prefix of new (tail windows slide)."""
    new = new_logs.splitlines()
    if not old_logs:
        return new
    old = old_logs.splitlines()
    overlap = 0
    for i in range(1, min(len(old), len(new)) + 1):
        if old[-i:] == new[:i]:
            overlap = i
    return new[overlap:]
