"""Tunnel failures. Every class carries a one-line ``hint`` that the CLI prints under the message
(same class names as the reference's hierarchy: packages/prime-tunnel/src/prime_tunnel/exceptions.py:1-48)."""

from __future__ import annotations


class TunnelError(Exception):
    """Root of everything ``prime_b200.platform.tunnel`` raises on purpose."""

    hint: str | None = None

    def __init__(self, message: str = "", *, hint: str | None = None) -> None:
        super().__init__(message)
        if hint is not None:
            self.hint = hint


def _kind(name: str, doc: str, hint: str) -> type[TunnelError]:
    return type(name, (TunnelError,), {"__doc__": doc, "hint": hint, "__module__": __name__})


BinaryDownloadError = _kind("BinaryDownloadError", "The pinned frpc release could not be fetched or failed its SHA-256 check.",
                            "check network access to github.com, or place a verified frpc under ~/.prime/bin")  # fmt: skip
TunnelAuthError = _kind("TunnelAuthError", "frps rejected the tunnel's token (expired registration or wrong account).",
                        "run `prime login` and start the tunnel again")  # fmt: skip
TunnelTimeoutError = _kind("TunnelTimeoutError", "frpc did not report a working proxy within the start-up window.",
                           "the local port may not be listening yet; raise the timeout or retry")  # fmt: skip
TunnelLimitReachedError = _kind("TunnelLimitReachedError", "The account already runs its maximum number of tunnels.",
                                "`prime tunnel list` / `prime tunnel stop <id>` frees a slot")  # fmt: skip


class TunnelConnectionError(TunnelError):
    """An operation needed a live tunnel and there is none (never started, stopped, or frpc died)."""

    hint = "start it with Tunnel.start() / `prime tunnel start <port>`"

    def __init__(self, message: str | None = None, *, tunnel_id: str | None = None) -> None:
        self.tunnel_id = tunnel_id
        if message is None:
            message = f"Tunnel {tunnel_id} is not running" if tunnel_id else "Tunnel is not running"
        super().__init__(message)
