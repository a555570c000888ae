"""Async REST client for ``/tunnel`` with two retry classes: registering a tunnel (POST) never retries a
time-out (the server may have created it); GET/DELETE also retry time-outs
(reference: packages/prime-tunnel/src/prime_tunnel/core/client.py:23-31, :93-133, :164-347)."""

from __future__ import annotations

import asyncio
from typing import Any

import httpx

from ..core.client import IDEMPOTENT_RETRY, TRANSPORT_RETRY, RetryPolicy
from ..core.client import user_agent as _default_user_agent
from ..core.config import Config
from .exceptions import TunnelAuthError, TunnelError, TunnelLimitReachedError, TunnelTimeoutError
from .models import TunnelInfo


def interpret(resp: httpx.Response, operation: str) -> dict[str, Any]:
    code = resp.status_code
    if code == 401:
        raise TunnelAuthError("API key unauthorized. Check PRIME_API_KEY.")
    if code == 402:
        raise TunnelAuthError("Payment required. Check billing status.")
    if code == 404:
        return {}
    if code >= 400:
        try:
            detail = str(resp.json().get("detail", resp.text))
        except Exception:
            detail = resp.text
        if code == 400 and "maximum number of" in detail.lower():
            raise TunnelLimitReachedError(detail)
        raise TunnelError(f"Failed to {operation}: {detail}")
    if code == 204 or not resp.content:
        return {}
    return resp.json()


class TunnelClient:
    def __init__(self, api_key: str | None = None, timeout: float = 30.0, config: Config | None = None,
                 transport: httpx.AsyncBaseTransport | None = None, user_agent: str | None = None):  # fmt: skip
        self.config = config or Config(writable=False)
        self.api_key = api_key or self.config.api_key
        self.base_url = self.config.base_url
        self._timeout, self._transport = timeout, transport
        self._headers = {"Content-Type": "application/json", "User-Agent": user_agent or _default_user_agent("prime-b200-tunnel")}
        if self.api_key:
            self._headers["Authorization"] = f"Bearer {self.api_key}"
        self._http: httpx.AsyncClient | None = None

    def _need_key(self) -> None:
        if not self.api_key:
            raise TunnelError("No API key configured. Set PRIME_API_KEY environment variable.")

    async def _client(self) -> httpx.AsyncClient:
        if self._http is None or self._http.is_closed:
            self._http = httpx.AsyncClient(timeout=self._timeout, headers=self._headers, follow_redirects=True, transport=self._transport)
        return self._http

    async def close(self) -> None:
        if self._http is not None and not self._http.is_closed:
            await self._http.aclose()
        self._http = None

    async def _send(self, policy: RetryPolicy, method: str, path: str, **kw: Any) -> httpx.Response:
        http = await self._client()
        attempt = 0
        while True:
            try:
                return await http.request(method, f"{self.base_url}/api/v1{path}", **kw)
            except (httpx.TimeoutException, TimeoutError) as e:
                attempt += 1
                if isinstance(e, httpx.TimeoutException) and attempt < policy.attempts and policy.should_retry(method, e):
                    await asyncio.sleep(policy.delay(attempt - 1))
                    continue
                raise TunnelTimeoutError(f"Request timed out: {e}") from e
            except httpx.RequestError as e:
                attempt += 1
                if attempt < policy.attempts and policy.should_retry(method, e):
                    await asyncio.sleep(policy.delay(attempt - 1))
                    continue
                raise TunnelError(f"Failed to connect to API: {e}") from e

    async def create_tunnel(self, local_port: int, name: str | None = None, team_id: str | None = None) -> TunnelInfo:
        self._need_key()
        body: dict[str, Any] = {"local_port": local_port}
        if name:
            body["name"] = name
        team_id = team_id if team_id is not None else self.config.team_id
        if team_id:
            body["teamId"] = team_id
        data = interpret(await self._send(TRANSPORT_RETRY, "POST", "/tunnel", json=body), "create tunnel")
        if not data:
            raise TunnelError("Failed to create tunnel: unexpected empty response")
        return TunnelInfo(**data)

    async def get_tunnel(self, tunnel_id: str) -> TunnelInfo | None:
        self._need_key()
        resp = await self._send(IDEMPOTENT_RETRY, "GET", f"/tunnel/{tunnel_id}")
        if resp.status_code == 404:
            return None
        return TunnelInfo.from_status(interpret(resp, "get tunnel"))

    async def delete_tunnel(self, tunnel_id: str) -> bool:
        self._need_key()
        resp = await self._send(IDEMPOTENT_RETRY, "DELETE", f"/tunnel/{tunnel_id}")
        if resp.status_code == 404:
            return False
        interpret(resp, "delete tunnel")
        return True

    async def bulk_delete_tunnels(self, tunnel_ids: list[str]) -> dict:
        self._need_key()
        return interpret(await self._send(IDEMPOTENT_RETRY, "DELETE", "/tunnel", json={"tunnel_ids": tunnel_ids}), "bulk delete tunnels")

    async def list_tunnels(self, team_id: str | None = None) -> list[TunnelInfo]:
        self._need_key()
        team_id = team_id if team_id is not None else self.config.team_id
        resp = await self._send(IDEMPOTENT_RETRY, "GET", "/tunnel", params={"teamId": team_id} if team_id else None)
        return [TunnelInfo.from_status(t) for t in interpret(resp, "list tunnels").get("tunnels", [])]

    async def __aenter__(self):
        return self

    async def __aexit__(self, *exc):
        await self.close()
