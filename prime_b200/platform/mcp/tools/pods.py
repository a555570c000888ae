"""(reference: packages/prime-mcp-server/src/prime_mcp/tools/pods.py:6-229)"""

from __future__ import annotations

from typing import Any

from ..client import call

_POSITIVE = ("gpu_count", "disk_size", "vcpus", "memory")
# python arg → wire key inside the "pod" object; second element: include when value is merely falsy?
_POD_OPTIONALS = [("name", "name", False), ("disk_size", "diskSize", True), ("vcpus", "vcpus", True), ("memory", "memory", True),
                  ("max_price", "maxPrice", True), ("country", "country", False), ("security", "security", False),
                  ("auto_restart", "autoRestart", True), ("jupyter_password", "jupyterPassword", False),
                  ("custom_template_id", "customTemplateId", False)]  # fmt: skip


def build_create_body(cloud_id: str, gpu_type: str, provider_type: str, data_center_id: str, gpu_count: int = 1,
                      socket: str = "PCIe", image: str = "ubuntu_22_cuda_12", env_vars: dict[str, str] | None = None,
                      team_id: str | None = None, **opt: Any) -> dict[str, Any]:  # fmt: skip
    """Nested ``{pod, provider, team}`` body; raises ValueError on non-positive sizes."""
    sizes = {"gpu_count": gpu_count, **{k: opt.get(k) for k in _POSITIVE[1:]}}
    for k, v in sizes.items():
        if k == "gpu_count" and v <= 0:
            raise ValueError("gpu_count must be greater than 0")
        if k != "gpu_count" and v is not None and v <= 0:
            raise ValueError(f"{k} must be greater than 0 if specified")
    pod: dict[str, Any] = {"cloudId": cloud_id, "gpuType": gpu_type, "gpuCount": gpu_count, "socket": socket, "image": image,
                           "dataCenterId": data_center_id}  # fmt: skip
    for arg, wire, keep_falsy in _POD_OPTIONALS:
        v = opt.get(arg)
        if v is None or (not keep_falsy and not v):
            continue
        pod[wire] = v
    if env_vars:
        pod["envVars"] = [{"key": k, "value": v} for k, v in env_vars.items()]
    body: dict[str, Any] = {"pod": pod, "provider": {"type": provider_type}}
    if team_id:
        body["team"] = {"teamId": team_id}
    return body


async def create_pod(cloud_id: str, gpu_type: str, provider_type: str, data_center_id: str, **kw: Any) -> dict[str, Any]:
    try:
        body = build_create_body(cloud_id, gpu_type, provider_type, data_center_id, **kw)
    except ValueError as e:
        return {"error": str(e)}
    return await call("POST", "pods/", "Unable to create pod", json_data=body)


async def list_pods(offset: int = 0, limit: int = 100) -> dict[str, Any]:
    return await call("GET", "pods/", "Unable to fetch pods list", params={"offset": max(0, offset), "limit": max(0, limit)})


async def get_pods_history(limit: int = 100, offset: int = 0, sort_by: str = "terminatedAt", sort_order: str = "desc") -> dict[str, Any]:
    params = {"limit": max(0, limit), "offset": max(0, offset), "sort_by": sort_by, "sort_order": sort_order}
    return await call("GET", "pods/history", "Unable to fetch pods history", params=params)


async def get_pods_status(pod_ids: list[str] | None = None) -> dict[str, Any]:
    return await call("GET", "pods/status", "Unable to fetch pods status", params={"pod_ids": pod_ids} if pod_ids else {})


async def get_pod_details(pod_id: str) -> dict[str, Any]:
    return await call("GET", f"pods/{pod_id}", f"Unable to fetch details for pod ID: {pod_id}")


async def delete_pod(pod_id: str) -> dict[str, Any]:
    return await call("DELETE", f"pods/{pod_id}", f"Unable to delete pod ID: {pod_id}")
