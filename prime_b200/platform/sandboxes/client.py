"""The sandbox SDK's own control-plane clients: the shared ``core.client`` classes with the SDK's defaults.

A program that builds ``prime_sandboxes.APIClient(api_key=...)`` itself (instead of letting ``SandboxClient`` do it) gets what the
reference SDK gives it — three attempts with 0.1 → 2 s exponential back-off on failures that prove the request never completed
(``RemoteProtocolError``, ``ConnectError``, ``PoolTimeout``, ``ReadError``) and the SDK's user agent — not the CLI's single shot
(reference: packages/prime-sandboxes/src/prime_sandboxes/core/client.py:19-24, 88-126, 199-237).
"""

from __future__ import annotations

from typing import Any

from ..core import client as _core
from ..core.client import (  # noqa: F401  (same names as ``core.client``: this module stands in for it under ``prime_sandboxes.core.client``)
    IDEMPOTENT_RETRY,
    NO_RETRY,
    TRANSPORT_ERRORS,
    TRANSPORT_RETRY,
    APIError,
    APITimeoutError,
    PaymentRequiredError,
    RetryPolicy,
    UnauthorizedError,
    ValidationError,
    user_agent,
)
from ..core.config import Config  # noqa: F401

RETRYABLE_EXCEPTIONS = TRANSPORT_ERRORS


def _agent() -> str:
    return user_agent("prime-b200-sandboxes")


class APIClient(_core.APIClient):
    def __init__(self, api_key: str | None = None, require_auth: bool = True, user_agent: str | None = None,
                 retry: RetryPolicy = TRANSPORT_RETRY, **kw: Any) -> None:  # fmt: skip
        super().__init__(api_key, require_auth, user_agent or _agent(), retry, **kw)


class AsyncAPIClient(_core.AsyncAPIClient):
    def __init__(self, api_key: str | None = None, require_auth: bool = True, user_agent: str | None = None,
                 retry: RetryPolicy = TRANSPORT_RETRY, **kw: Any) -> None:  # fmt: skip
        super().__init__(api_key, require_auth, user_agent or _agent(), retry, **kw)
