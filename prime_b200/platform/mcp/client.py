"""Error-swallowing request shim for MCP tools: a tool call must always return a dict, so every failure
becomes ``{"error": "..."}`` instead of an exception (reference: packages/prime-mcp-server/src/prime_mcp/client.py:8-37).
The shared AsyncAPIClient is created lazily (importing the server must not need credentials)."""

from __future__ import annotations

from typing import Any

from ..core.client import AsyncAPIClient, user_agent

_client: AsyncAPIClient | None = None


def get_client() -> AsyncAPIClient:
    global _client
    if _client is None:
        _client = AsyncAPIClient(user_agent=user_agent("prime-b200-mcp"))
    return _client


def set_client(client: AsyncAPIClient | None) -> None:
    """Test hook / embedding hook."""
    global _client
    _client = client


_VERBS = {"GET", "POST", "DELETE", "PATCH"}


async def make_prime_request(method: str, endpoint: str, params: dict[str, Any] | None = None,
                             json_data: dict[str, Any] | None = None) -> dict[str, Any]:  # fmt: skip
    if method not in _VERBS:
        return {"error": f"Unsupported HTTP method: {method}"}
    try:
        c = get_client()
        if method == "GET":
            return await c.get(endpoint, params=params)
        if method == "POST":
            return await c.post(endpoint, json=json_data)
        if method == "PATCH":
            return await c.patch(endpoint, json=json_data)
        return await c.delete(endpoint)
    except Exception as e:
        return {"error": str(e)}


async def call(method: str, endpoint: str, failure: str, **kw: Any) -> dict[str, Any]:
    data = await make_prime_request(method, endpoint, **kw)
    return data if data else {"error": failure}
