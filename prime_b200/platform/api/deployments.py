"""LoRA adapter deployments (reference: packages/prime/src/prime_cli/api/deployments.py:35-103;
endpoints /rft/adapters[/{id}[/deploy|/unload]], /rft/deployable-models)."""

from __future__ import annotations

from datetime import datetime
from typing import Any

from ._base import ApiModel, wrap

DEPLOYABLE_FROM = {"NOT_DEPLOYED", "DEPLOY_FAILED", "UNLOAD_FAILED"}
UNLOADABLE_FROM = {"DEPLOYED", "DEPLOY_FAILED", "UNLOAD_FAILED"}


class Adapter(ApiModel):
    id: str
    display_name: str | None = None
    user_id: str
    team_id: str | None = None
    rft_run_id: str
    base_model: str
    step: int | None = None
    status: str  # PENDING | UPLOADING | READY | FAILED
    deployment_status: str = "NOT_DEPLOYED"
    deployed_at: datetime | None = None
    deployment_error: str | None = None
    created_at: datetime
    updated_at: datetime


class DeploymentsClient:
    def __init__(self, client: Any) -> None:
        self.client = client

    def list_adapters(self, team_id: str | None = None, limit: int | None = None, offset: int = 0) -> tuple[list[Adapter], int]:
        params = {k: v for k, v in (("team_id", team_id), ("limit", limit), ("offset", offset or None)) if v is not None}
        with wrap("list adapters"):
            resp = self.client.get("/rft/adapters", params=params or None)
            rows = resp.get("adapters", [])
            return [Adapter.model_validate(a) for a in rows], resp.get("total", len(rows))

    def get_adapter(self, adapter_id: str) -> Adapter:
        with wrap("get adapter"):
            return Adapter.model_validate(self.client.get(f"/rft/adapters/{adapter_id}").get("adapter"))

    def deploy_adapter(self, adapter_id: str) -> Adapter:
        with wrap("deploy adapter"):
            return Adapter.model_validate(self.client.post(f"/rft/adapters/{adapter_id}/deploy").get("adapter"))

    def unload_adapter(self, adapter_id: str) -> Adapter:
        with wrap("unload adapter"):
            return Adapter.model_validate(self.client.post(f"/rft/adapters/{adapter_id}/unload").get("adapter"))

    def get_deployable_models(self) -> list[str]:
        with wrap("get deployable models"):
            return self.client.get("/rft/deployable-models").get("models") or []
