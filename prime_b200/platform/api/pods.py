"""Pods: list / status / get / create / delete / history
(reference: packages/prime/src/prime_cli/api/pods.py:21-240; endpoints /pods, /pods/status, /pods/{id}, /pods/history)."""

from __future__ import annotations

from typing import Any

from pydantic import Field, field_validator

from ._base import ApiModel, unwrap_single_none, wrap

clean_connection_fields = unwrap_single_none  # the reference's public name of the same validator helper (api/pods.py:9)


class PortMapping(ApiModel):
    internal: str
    external: str
    protocol: str
    used_by: str | None = None
    description: str | None = None


class _Connectable(ApiModel):
    # multi-node pods report one entry per node; nodes that are still provisioning come back as null
    ssh_connection: str | list[str | None] | None = None
    ip: str | list[str | None] | None = None

    @field_validator("ssh_connection", "ip", mode="before")
    @classmethod
    def _clean(cls, v: Any) -> Any:
        return unwrap_single_none(v)


class PodStatus(_Connectable):
    pod_id: str
    provider_type: str
    status: str
    cost_per_hr: float | None = Field(None, alias="priceHr")
    prime_port_mapping: list[PortMapping] | None = None
    installation_failure: str | None = None
    installation_progress: int | None = None


class AttachedResource(ApiModel):
    id: str | int | None = None
    type: str | None = None
    status: str | None = None
    created_at: str | None = None
    size: int | None = None
    mount_path: str | None = None
    resource_path: str | None = None
    is_detachable: bool | None = None
    resource_type: str | None = None


class Pod(_Connectable):
    id: str
    name: str | None = None
    gpu_type: str = Field(..., alias="gpuName")
    gpu_count: int
    status: str
    created_at: str
    provider_type: str
    installation_status: str | None = None
    installation_failure: str | None = None
    installation_progress: int | None = None
    team_id: str | None = None
    resources: dict | None = None
    attached_resources: list[AttachedResource] | None = None
    prime_port_mapping: list[PortMapping] | None = None
    price_hr: float | None = None
    environment_type: str | None = None
    socket: str | None = None
    type: str | None = None
    user_id: str | None = None
    wallet_id: str | None = None
    updated_at: str | None = None
    jupyter_password: str | None = None
    stopped_price_hr: float | None = None
    provisioning_price_hr: float | None = None
    base_price_hr: float | None = None
    base_currency: str | None = None
    custom_template_id: str | None = None
    is_spot: bool | None = None
    auto_restart: bool | None = None


class PodList(ApiModel):
    total_count: int = Field(..., alias="total_count")
    offset: int
    limit: int
    data: list[Pod]


class PodConfig(ApiModel):
    name: str | None = None
    cloud_id: str
    gpu_type: str
    socket: str
    gpu_count: int
    disk_size: int | None = None
    vcpus: int | None = None
    memory: int | None = None
    image: str | None = None
    custom_template_id: str | None = None
    data_center_id: str | None = None
    country: str | None = None
    security: str | None = None
    auto_restart: bool | None = None
    provider: dict
    team: dict | None = None


class HistoryObj(ApiModel):
    id: str
    name: str
    provider_type: str
    provisioned_by: str | None = None
    type: str
    created_at: str
    terminated_at: str | None = None
    gpu_name: str
    count: int = Field(..., alias="gpuCount")
    socket: str | None = None
    price_hr: float
    user_id: str
    team_id: str | None = None
    total_billed_price: float


class HistoryList(ApiModel):
    total_count: int = Field(..., alias="total_count")
    offset: int
    limit: int
    data: list[HistoryObj]


class PodsClient:
    def __init__(self, client: Any) -> None:
        self.client = client

    def list(self, offset: int = 0, limit: int = 100) -> PodList:
        with wrap("list pods"):
            return PodList.model_validate(self.client.get("/pods", params={"offset": offset, "limit": limit}))

    def get_status(self, pod_ids: list[str]) -> list[PodStatus]:
        with wrap("get pod status"):
            resp = self.client.get("/pods/status", params={"pod_ids": pod_ids})
            return [PodStatus.model_validate(s) for s in resp.get("data", [])]

    def get(self, pod_id: str) -> Pod:
        with wrap("get pod details"):
            return Pod.model_validate(self.client.get(f"/pods/{pod_id}"))

    def create(self, pod_config: dict) -> Pod:
        team_id = getattr(getattr(self.client, "config", None), "team_id", None)
        if not pod_config.get("team") and team_id:
            pod_config = {**pod_config, "team": {"teamId": team_id}}
        with wrap("create pod"):
            return Pod.model_validate(self.client.post("/pods", json=pod_config))

    def delete(self, pod_id: str) -> None:
        with wrap("delete pod"):
            self.client.delete(f"/pods/{pod_id}")

    def history(self, offset: int = 0, limit: int = 100) -> HistoryList:
        with wrap("get pods history"):
            return HistoryList.model_validate(self.client.get("/pods/history", params={"offset": offset, "limit": limit}))
