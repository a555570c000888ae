"""(reference: packages/prime-tunnel/src/prime_tunnel/exceptions.py:1-48)"""


class TunnelError(Exception):
    pass


class TunnelConnectionError(TunnelError):
    def __init__(self, message: str | None = None, *, tunnel_id: str | None = None):
        self.tunnel_id = tunnel_id
        super().__init__(message or (f"Tunnel {tunnel_id} is not running" if tunnel_id else "Tunnel is not running"))


class TunnelAuthError(TunnelError):
    pass


class TunnelTimeoutError(TunnelError):
    pass


class TunnelLimitReachedError(TunnelError):
    pass


class BinaryDownloadError(TunnelError):
    pass
