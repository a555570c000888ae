"""HTTP transport shared by every SDK: sync + async clients over httpx.

Parity targets (behaviour, not code):
  * reference packages/prime/src/prime_cli/core/client.py:58-342 — Bearer auth, ``/api/v1`` prefixing,
    204 → ``{}``, non-dict JSON → error, 401/402/422/timeout/transport → typed exceptions;
  * reference packages/prime-sandboxes/src/prime_sandboxes/core/client.py:19-103 — optional retry of
    connection-level failures (3 attempts, exponential back-off 0.1–2 s).

The retry policy is a small explicit object (``RetryPolicy``) instead of tenacity decorators so the same
client serves the CLI (no retry), the sandbox SDK (transport retry) and the tunnel SDK (idempotent verbs
also retry time-outs; reference packages/prime-tunnel/src/prime_tunnel/core/client.py:23-31, :93-133).
"""

from __future__ import annotations

import asyncio
import random
import sys
import time
from dataclasses import dataclass, field
from typing import Any

import httpx

from .config import Config


class APIError(Exception):
    """Base class for every platform API failure."""

    def __init__(self, message: str, status_code: int | None = None):
        super().__init__(message)
        self.status_code = status_code


class UnauthorizedError(APIError):
    pass


class PaymentRequiredError(APIError):
    pass


class APITimeoutError(APIError):
    pass


class ValidationError(APIError):
    """HTTP 422 with a FastAPI-style ``detail`` list."""

    def __init__(self, errors: list[dict]):
        self.errors = errors
        rows = ["Validation failed:"]
        for e in errors:
            where = ".".join(str(p) for p in e.get("loc", []) if p != "body")
            rows.append(f"  - {where}: {e.get('msg', 'unknown error')}")
        super().__init__("\n".join(rows), 422)


TRANSPORT_ERRORS = (httpx.RemoteProtocolError, httpx.ConnectError, httpx.PoolTimeout, httpx.ReadError)


@dataclass
class RetryPolicy:
    attempts: int = 1
    base_delay: float = 0.1
    max_delay: float = 2.0
    retry_timeouts_for: frozenset[str] = field(default_factory=frozenset)  # HTTP verbs whose time-outs are retried
    retry_statuses: frozenset[int] = field(default_factory=frozenset)

    def should_retry(self, method: str, exc: BaseException) -> bool:
        if isinstance(exc, TRANSPORT_ERRORS):
            return True
        if isinstance(exc, httpx.TimeoutException):
            return method.upper() in self.retry_timeouts_for
        if isinstance(exc, httpx.HTTPStatusError):
            return exc.response.status_code in self.retry_statuses
        return False

    def delay(self, attempt: int) -> float:
        d = min(self.max_delay, self.base_delay * (2**attempt))
        return d * (0.5 + random.random() / 2)


NO_RETRY = RetryPolicy(attempts=1)
TRANSPORT_RETRY = RetryPolicy(attempts=3, base_delay=0.1, max_delay=2.0)
IDEMPOTENT_RETRY = RetryPolicy(attempts=3, base_delay=0.2, max_delay=4.0, retry_timeouts_for=frozenset({"GET", "DELETE"}))


def user_agent(product: str = "prime-b200-cli") -> str:
    from .. import __version__

    v = sys.version_info
    return f"{product}/{__version__} python/{v.major}.{v.minor}.{v.micro}"


def _api_path(endpoint: str) -> str:
    return f"/api/v1{endpoint}" if endpoint.startswith("/") else f"/api/v1/{endpoint}"


def _raise_for(e: Exception) -> None:
    """Translate an httpx exception into the typed hierarchy above."""
    if isinstance(e, httpx.HTTPStatusError):
        code = e.response.status_code
        if code == 401:
            raise UnauthorizedError(
                "API key unauthorized. Check the key's permissions, create a new one at "
                "https://app.primeintellect.ai/dashboard/tokens or run 'prime login'.",
                401,
            ) from e
        if code == 402:
            raise PaymentRequiredError(
                "Payment required. Check your billing status at https://app.primeintellect.ai/dashboard/billing", 402
            ) from e
        body: Any = None
        try:
            body = e.response.json()
        except ValueError:
            pass
        if code == 422 and isinstance(body, dict) and isinstance(body.get("detail"), list):
            raise ValidationError(body["detail"]) from e
        if isinstance(body, dict) and "detail" in body:
            raise APIError(f"HTTP {code}: {body['detail']}", code) from e
        raise APIError(f"HTTP {code}: {e.response.text or e}", code) from e
    if isinstance(e, httpx.TimeoutException):
        raise APITimeoutError(f"Request timed out: {e}") from e
    if isinstance(e, httpx.RequestError):
        req = getattr(e, "_request", None)
        where = f"{req.method} {req.url}" if req is not None else "?"
        raise APIError(f"Request failed: {type(e).__name__} at {where}: {e}") from e
    raise e


def _decode(resp: httpx.Response) -> dict[str, Any]:
    if resp.status_code == 204 and not resp.content:
        return {}
    data = resp.json()
    if not isinstance(data, dict):
        raise APIError("API response was not a dictionary")
    return data


class _Base:
    def __init__(self, api_key: str | None, require_auth: bool, agent: str | None, retry: RetryPolicy,
                 config: Config | None):  # fmt: skip
        self.config = config or Config(writable=False)
        self.api_key = api_key or self.config.api_key
        self.require_auth = require_auth
        self.base_url = self.config.base_url
        self.retry = retry
        self.headers = {"Content-Type": "application/json", "User-Agent": agent or user_agent()}
        if self.api_key:
            self.headers["Authorization"] = f"Bearer {self.api_key}"

    def _check_auth(self) -> None:
        if self.require_auth and not self.api_key:
            raise APIError("No API key configured. Run 'prime login' (or set PRIME_API_KEY).")

    def __repr__(self) -> str:
        return f"{type(self).__name__}(base_url={self.base_url})"


class APIClient(_Base):
    def __init__(self, api_key: str | None = None, require_auth: bool = True, user_agent: str | None = None,
                 retry: RetryPolicy = NO_RETRY, config: Config | None = None, transport: httpx.BaseTransport | None = None):  # fmt: skip
        super().__init__(api_key, require_auth, user_agent, retry, config)
        self.client = httpx.Client(headers=self.headers, follow_redirects=True,
                                   timeout=httpx.Timeout(30.0, connect=10.0), transport=transport)  # fmt: skip

    def request(self, method: str, endpoint: str, params: dict | None = None, json: Any = None,
                timeout: float | None = None) -> dict[str, Any]:  # fmt: skip
        self._check_auth()
        url = self.base_url + _api_path(endpoint)
        kwargs: dict[str, Any] = {"params": params, "json": json}
        if timeout is not None:
            kwargs["timeout"] = timeout
        attempt = 0
        while True:
            try:
                resp = self.client.request(method, url, **kwargs)
                resp.raise_for_status()
                return _decode(resp)
            except (httpx.HTTPError, httpx.TimeoutException) as e:
                attempt += 1
                if attempt < self.retry.attempts and self.retry.should_retry(method, e):
                    time.sleep(self.retry.delay(attempt - 1))
                    continue
                _raise_for(e)

    def get(self, endpoint: str, params: dict | None = None) -> dict[str, Any]:
        return self.request("GET", endpoint, params=params)

    def post(self, endpoint: str, json: Any = None, params: dict | None = None) -> dict[str, Any]:
        return self.request("POST", endpoint, json=json, params=params)

    def put(self, endpoint: str, json: Any = None) -> dict[str, Any]:
        return self.request("PUT", endpoint, json=json)

    def patch(self, endpoint: str, json: Any = None, params: dict | None = None) -> dict[str, Any]:
        return self.request("PATCH", endpoint, json=json, params=params)

    def delete(self, endpoint: str, params: dict | None = None, json: Any = None) -> dict[str, Any]:
        return self.request("DELETE", endpoint, params=params, json=json)

    def close(self) -> None:
        self.client.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class AsyncAPIClient(_Base):
    def __init__(self, api_key: str | None = None, require_auth: bool = True, user_agent: str | None = None,
                 retry: RetryPolicy = NO_RETRY, config: Config | None = None,
                 transport: httpx.AsyncBaseTransport | None = None, limits: httpx.Limits | None = None):  # fmt: skip
        super().__init__(api_key, require_auth, user_agent, retry, config)
        self.client = httpx.AsyncClient(headers=self.headers, follow_redirects=True,
                                        timeout=httpx.Timeout(30.0, connect=10.0), transport=transport,
                                        limits=limits or httpx.Limits())  # fmt: skip

    async def request(self, method: str, endpoint: str, params: dict | None = None, json: Any = None,
                      timeout: float | None = None) -> dict[str, Any]:  # fmt: skip
        self._check_auth()
        url = self.base_url + _api_path(endpoint)
        kwargs: dict[str, Any] = {"params": params, "json": json}
        if timeout is not None:
            kwargs["timeout"] = timeout
        attempt = 0
        while True:
            try:
                resp = await self.client.request(method, url, **kwargs)
                resp.raise_for_status()
                return _decode(resp)
            except (httpx.HTTPError, httpx.TimeoutException) as e:
                attempt += 1
                if attempt < self.retry.attempts and self.retry.should_retry(method, e):
                    await asyncio.sleep(self.retry.delay(attempt - 1))
                    continue
                _raise_for(e)

    async def get(self, endpoint: str, params: dict | None = None) -> dict[str, Any]:
        return await self.request("GET", endpoint, params=params)

    async def post(self, endpoint: str, json: Any = None, params: dict | None = None) -> dict[str, Any]:
        return await self.request("POST", endpoint, json=json, params=params)

    async def put(self, endpoint: str, json: Any = None) -> dict[str, Any]:
        return await self.request("PUT", endpoint, json=json)

    async def patch(self, endpoint: str, json: Any = None, params: dict | None = None) -> dict[str, Any]:
        return await self.request("PATCH", endpoint, json=json, params=params)

    async def delete(self, endpoint: str, params: dict | None = None, json: Any = None) -> dict[str, Any]:
        return await self.request("DELETE", endpoint, params=params, json=json)

    async def aclose(self) -> None:
        await self.client.aclose()

    async def __aenter__(self):
        return self

    async def __aexit__(self, *exc):
        await self.aclose()
