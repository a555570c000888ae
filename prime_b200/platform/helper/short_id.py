"""Stable 6-hex short ids for GPU / disk offers: md5 of the identifying tuple
(reference: packages/prime/src/prime_cli/helper/short_id.py:6-32 — same tuple, so ids match)."""

from __future__ import annotations

import hashlib

from ..api.availability import DiskAvailability, GPUAvailability


def _digest(parts: list) -> str:
    return hashlib.md5("-".join(str(p) for p in parts).encode()).hexdigest()[:6]


def generate_short_id(g: GPUAvailability) -> str:
    na = "N/A"
    location = f"{g.country or na} - {g.data_center or na}"
    return _digest([g.cloud_id, g.gpu_type, g.socket or na, location, g.provider or na,
                    g.memory.default_count, g.vcpu.default_count, g.gpu_count])  # fmt: skip


def generate_short_id_disk(d: DiskAvailability) -> str:
    return _digest([d.provider, d.cloud_id, d.data_center, d.country, d.region, d.spec.default_count, ""])
