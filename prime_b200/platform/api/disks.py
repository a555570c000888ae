"""Persistent disks: list/get/create/update/delete
(reference: packages/prime/src/prime_cli/api/disks.py:71-150; endpoints /disks, /disks/{id})."""

from __future__ import annotations

from typing import Any

from pydantic import Field

from ._base import ApiModel, wrap


class DiskInfo(ApiModel):
    country: str | None = None
    data_center_id: str | None = None
    cloud_id: str | None = None
    is_multinode: bool | None = None


class Disk(ApiModel):
    id: str
    name: str
    created_at: str
    updated_at: str
    terminated_at: str | None = None
    status: str
    provider_type: str
    size: int
    info: dict | None = None
    price_hr: float | None = None
    stopped_price_hr: float | None = None
    provisioning_price_hr: float | None = None
    user_id: str | None = None
    team_id: str | None = None
    wallet_id: str | None = None
    pods: list[str] = Field(default_factory=list)
    clusters: list[str] = Field(default_factory=list)


class DiskList(ApiModel):
    total_count: int = Field(..., alias="total_count")
    offset: int
    limit: int
    data: list[Disk]


class DiskCreateRequest(ApiModel):
    size: int = Field(..., gt=0)
    name: str | None = None
    country: str | None = None
    cloud_id: str | None = None
    data_center_id: str | None = None


class DiskUpdateRequest(ApiModel):
    """Body of ``PATCH /disks/{id}`` (reference: packages/prime/src/prime_cli/api/disks.py, DiskUpdateRequest)."""

    name: str


class DiskDeleteResponse(ApiModel):
    status: str


class DisksClient:
    def __init__(self, client: Any) -> None:
        self.client = client

    def list(self, offset: int = 0, limit: int = 100) -> DiskList:
        with wrap("list disks"):
            return DiskList.model_validate(self.client.get("/disks", params={"offset": offset, "limit": limit}))

    def get(self, disk_id: str) -> Disk:
        with wrap("get disk details"):
            return Disk.model_validate(self.client.get(f"/disks/{disk_id}"))

    def create(self, disk_config: dict) -> Disk:
        team_id = getattr(getattr(self.client, "config", None), "team_id", None)
        if not disk_config.get("team") and team_id:
            disk_config = {**disk_config, "team": {"teamId": team_id}}
        with wrap("create disk"):
            return Disk.model_validate(self.client.post("/disks", json=disk_config))

    def update(self, disk_id: str, name: str) -> dict:
        with wrap("update disk"):
            return self.client.patch(f"/disks/{disk_id}", json={"name": name})

    def delete(self, disk_id: str) -> DiskDeleteResponse:
        with wrap("delete disk"):
            return DiskDeleteResponse.model_validate(self.client.delete(f"/disks/{disk_id}"))
