"""(reference: packages/prime-tunnel/src/prime_tunnel/models.py:7-23)"""

from __future__ import annotations

from datetime import datetime

from pydantic import BaseModel, ConfigDict


class TunnelInfo(BaseModel):
    model_config = ConfigDict(from_attributes=True)
    tunnel_id: str
    hostname: str
    url: str
    frp_token: str = ""  # only returned by create
    binding_secret: str = ""
    server_host: str = ""
    server_port: int = 7000
    expires_at: datetime
    user_id: str | None = None
    status: str | None = None

    @classmethod
    def from_status(cls, d: dict) -> "TunnelInfo":
        """Status/list responses carry no credentials."""
        return cls(tunnel_id=d["tunnel_id"], hostname=d["hostname"], url=d["url"], expires_at=d["expires_at"],
                   user_id=d.get("user_id"), status=d.get("status"))  # fmt: skip
