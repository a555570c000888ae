"""Small pure formatters (reference: packages/prime/src/prime_cli/utils/formatters.py:4-70)."""

from __future__ import annotations

import re
from typing import Any, Mapping

_ANSI = re.compile(r"\x1b\[[0-?]*[ -/]*[@-~]|\x1b\][^\x07]*\x07")


def strip_ansi(text: str) -> str:
    return _ANSI.sub("", text)


def obfuscate_secret(value: str | None, keep: int = 4) -> str:
    if not value:
        return ""
    if len(value) <= keep * 2:
        return "*" * len(value)
    return f"{value[:keep]}{'*' * (len(value) - 2 * keep)}{value[-keep:]}"


def obfuscate_env_vars(env_vars: Mapping[str, str] | None) -> dict[str, str]:
    return {k: obfuscate_secret(v, keep=2) for k, v in (env_vars or {}).items()}


def format_price(value: float | None, unit: str = "/hr") -> str:
    return "N/A" if value is None else f"${value:,.2f}{unit}"


def format_ip_display(ip: Any) -> str:
    if ip is None:
        return "N/A"
    if isinstance(ip, (list, tuple)):
        return ", ".join(str(i) for i in ip) if ip else "N/A"
    return str(ip)


def format_size(num_bytes: float | None) -> str:
    if num_bytes is None:
        return "N/A"
    n = float(num_bytes)
    for unit in ("B", "KB", "MB", "GB", "TB"):
        if abs(n) < 1024 or unit == "TB":
            return f"{n:.0f} {unit}" if unit == "B" else f"{n:.1f} {unit}"
        n /= 1024
    return f"{n:.1f} TB"


def format_resources(cpu_cores: float | None = None, memory_gb: float | None = None, gpu_count: int | None = 0) -> str:
    """Compact form used in list rows and in ``--output json`` (``2CPU/4GB``, ``8CPU/64GB/1GPU``; reference:
    packages/prime/src/prime_cli/utils/formatters.py:48-53 — scripts parse it)."""
    out = f"{(cpu_cores or 0):g}CPU/{(memory_gb or 0):g}GB"
    return out + (f"/{gpu_count}GPU" if gpu_count and gpu_count > 0 else "")


def format_resources_long(cpu: float | None = None, memory_gb: float | None = None, disk_gb: float | None = None,
                          gpu: int | None = None, gpu_type: str | None = None) -> str:  # fmt: skip
    """Spelled-out form for detail views: ``2 CPU, 4 GB RAM, 10 GB disk, 1x H100_80GB``."""
    parts = []
    if cpu is not None:
        parts.append(f"{cpu:g} CPU")
    if memory_gb is not None:
        parts.append(f"{memory_gb:g} GB RAM")
    if disk_gb is not None:
        parts.append(f"{disk_gb:g} GB disk")
    if gpu:
        parts.append(f"{gpu}x {gpu_type or 'GPU'}")
    return ", ".join(parts) or "N/A"


# ---- names the reference exports (packages/prime/src/prime_cli/utils/formatters.py:24-70), kept for drop-in imports
def obfuscate_secrets(secrets: Mapping[str, Any] | None) -> dict[str, str]:
    """Keys only: every value is shown as ``***`` (secrets are never echoed, not even partially)."""
    return {k: "***" for k in (secrets or {})}


def format_gpu_spec(gpu_type: str, gpu_count: int) -> str:
    return f"{gpu_type} x{gpu_count}"


def format_file_size(size_bytes: int) -> str:
    """``1536`` → ``1.5 KB``; plain byte counts below 1 KiB."""
    n = int(size_bytes)
    return f"{n} bytes" if n < 1024 else format_size(n)

