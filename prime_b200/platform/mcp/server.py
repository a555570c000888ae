"""The MCP server: ``FastMCP("primeintellect")`` with nine tools over stdio
(reference: packages/prime-mcp-server/src/prime_mcp/mcp.py:5-257).  Tool docstrings are the prompts an
agent sees, so they carry the operational warnings (SSH key first, spot vs on-demand, pick an image).

Run with ``python -m prime_b200.platform.mcp.server``."""

from __future__ import annotations

from mcp.server.fastmcp import FastMCP

from .tools import availability, pods, ssh

mcp = FastMCP("primeintellect")

_REGIONS = ('"africa", "asia_south", "asia_northeast", "australia", "canada", "eu_east", "eu_north", "eu_west", '
            '"middle_east", "south_america", "united_states"')  # fmt: skip
_SPOT_NOTE = ("Every offer has an 'isSpot' field: spot (true) is 50-90% cheaper but CAN BE TERMINATED AT ANY TIME; "
              "on-demand (false/null) costs more but is guaranteed. Show both prices and let the user decide; also show "
              "the available 'images'.")  # fmt: skip


@mcp.tool()
async def check_gpu_availability(gpu_type: str | None = None, regions: list[str] | None = None, socket: str | None = None,
                                 security: str | None = None, gpu_count: int | None = None) -> dict:  # fmt: skip
    return await availability.check_gpu_availability(gpu_type, regions, socket, security, gpu_count)


check_gpu_availability.__doc__ = f"""Check single-node GPU availability across providers.

gpu_type e.g. "H100_80GB", "A100_80GB", "B200_180GB"; regions from {_REGIONS}; socket one of "PCIe", "SXM2".."SXM6";
security "secure_cloud" | "community_cloud". Returns offers grouped by GPU type. {_SPOT_NOTE}"""


@mcp.tool()
async def check_cluster_availability(regions: list[str] | None = None, gpu_count: int | None = None, gpu_type: str | None = None,
                                     socket: str | None = None, security: str | None = None) -> dict:  # fmt: skip
    """Check multi-node cluster availability (same filters as check_gpu_availability); grouped by GPU type."""
    return await availability.check_cluster_availability(regions, gpu_count, gpu_type, socket, security)


@mcp.tool()
async def create_pod(cloud_id: str, gpu_type: str, provider_type: str, data_center_id: str, name: str | None = None,
                     gpu_count: int = 1, socket: str = "PCIe", disk_size: int | None = None, vcpus: int | None = None,
                     memory: int | None = None, max_price: float | None = None, image: str = "ubuntu_22_cuda_12",
                     custom_template_id: str | None = None, country: str | None = None, security: str | None = None,
                     auto_restart: bool | None = None, jupyter_password: str | None = None,
                     env_vars: dict[str, str] | None = None, team_id: str | None = None) -> dict:  # fmt: skip
    """Create a GPU pod.

    BEFORE CREATING: (1) the user must have an SSH key (manage_ssh_keys) or the pod is unreachable; (2) confirm spot
    vs on-demand from the offer's 'isSpot'; (3) ask which image they need (see the offer's 'images').
    cloud_id and data_center_id ('dataCenter' in availability results, e.g. "US-CA-2") come from an availability check;
    provider_type e.g. "runpod", "hyperstack", "datacrunch". Sizes must be positive when given."""
    return await pods.create_pod(cloud_id, gpu_type, provider_type, data_center_id, name=name, gpu_count=gpu_count, socket=socket,
                                 disk_size=disk_size, vcpus=vcpus, memory=memory, max_price=max_price, image=image,
                                 custom_template_id=custom_template_id, country=country, security=security,
                                 auto_restart=auto_restart, jupyter_password=jupyter_password, env_vars=env_vars, team_id=team_id)  # fmt: skip


@mcp.tool()
async def list_pods(offset: int = 0, limit: int = 100) -> dict:
    """List the user's pods (paginated)."""
    return await pods.list_pods(offset, limit)


@mcp.tool()
async def get_pods_history(limit: int = 100, offset: int = 0, sort_by: str = "terminatedAt", sort_order: str = "desc") -> dict:
    """Terminated-pod history; sort_by "terminatedAt" | "createdAt", sort_order "asc" | "desc"."""
    return await pods.get_pods_history(limit, offset, sort_by, sort_order)


@mcp.tool()
async def get_pods_status(pod_ids: list[str] | None = None) -> dict:
    """Status (state, ssh connection, install progress) of all pods or of the given ids."""
    return await pods.get_pods_status(pod_ids)


@mcp.tool()
async def get_pod_details(pod_id: str) -> dict:
    """Full details of one pod."""
    return await pods.get_pod_details(pod_id)


@mcp.tool()
async def delete_pod(pod_id: str) -> dict:
    """Terminate a pod. Irreversible: confirm with the user first."""
    return await pods.delete_pod(pod_id)


@mcp.tool()
async def manage_ssh_keys(action: str = "list", key_name: str | None = None, public_key: str | None = None,
                          key_id: str | None = None, offset: int = 0, limit: int = 100) -> dict:  # fmt: skip
    """Manage SSH keys: action "list" | "add" (key_name + public_key) | "delete" (key_id) | "set_primary" (key_id).
    A key MUST exist before creating pods; check with action="list" first."""
    return await ssh.manage_ssh_keys(action, key_name, public_key, key_id, offset, limit)


TOOLS = ("check_gpu_availability", "check_cluster_availability", "create_pod", "list_pods", "get_pods_history",
         "get_pods_status", "get_pod_details", "delete_pod", "manage_ssh_keys")  # fmt: skip


def main() -> None:
    mcp.run(transport="stdio")


if __name__ == "__main__":
    main()
