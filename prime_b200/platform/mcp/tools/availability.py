"""(reference: packages/prime-mcp-server/src/prime_mcp/tools/availability.py:6-91)"""

from __future__ import annotations

from typing import Any

from ..client import call


def _filters(**kw: Any) -> dict[str, Any]:
    return {k: v for k, v in kw.items() if v}


async def check_gpu_availability(gpu_type: str | None = None, regions: list[str] | None = None, socket: str | None = None,
                                 security: str | None = None, gpu_count: int | None = None) -> dict[str, Any]:  # fmt: skip
    params = _filters(regions=regions, gpu_type=gpu_type, socket=socket, security=security, gpu_count=gpu_count)
    return await call("GET", "availability/", "Unable to fetch GPU availability", params=params)


async def check_cluster_availability(regions: list[str] | None = None, gpu_count: int | None = None, gpu_type: str | None = None,
                                     socket: str | None = None, security: str | None = None) -> dict[str, Any]:  # fmt: skip
    params = _filters(regions=regions, gpu_count=gpu_count, gpu_type=gpu_type, socket=socket, security=security)
    return await call("GET", "availability/clusters", "Unable to fetch cluster availability", params=params)
