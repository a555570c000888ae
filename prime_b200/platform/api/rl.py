"""Hosted RL (``/rft``) client: models, runs CRUD, stop/restart, checkpoints, logs, metrics, rollouts,
progress, distributions, environment status
(reference: packages/prime/src/prime_cli/api/rl.py:20-423)."""

from __future__ import annotations

from datetime import datetime
from typing import Any

from pydantic import Field

from ..core.client import ValidationError
from ._base import ApiModel, wrap


class RLModel(ApiModel):
    name: str
    at_capacity: bool = False


class RLRun(ApiModel):
    id: str
    name: str | None = None
    user_id: str
    team_id: str | None = None
    cluster_id: str | None = Field(None, alias="rftClusterId")
    status: str
    rollouts_per_example: int
    seq_len: int
    max_steps: int
    max_tokens: int | None = None
    batch_size: int
    base_model: str
    environments: list[dict[str, Any]] = Field(default_factory=list)
    run_config: dict[str, Any] | None = None
    eval_config: dict[str, Any] | None = None
    val_config: dict[str, Any] | None = None
    buffer_config: dict[str, Any] | None = None
    learning_rate: float | None = None
    lora_alpha: int | None = None
    oversampling_factor: float | None = None
    max_async_level: int | None = None
    wandb_entity: str | None = None
    wandb_project: str | None = None
    wandb_run_name: str | None = None
    runs_ahead: int | None = None
    started_at: datetime | None = None
    completed_at: datetime | None = None
    error_message: str | None = None
    created_at: datetime
    updated_at: datetime


class RLCheckpoint(ApiModel):
    id: str
    rft_run_id: str
    step: int
    storage_url: str
    status: str
    size_bytes: int | None = None
    created_at: datetime
    uploaded_at: datetime | None = None


# (python keyword, wire key, include when falsy-but-not-None?)
_OPTIONAL_PAYLOAD = [
    ("name", "name", False), ("team_id", "team_id", False), ("max_tokens", "max_tokens", False),
    ("temperature", "temperature", True), ("repetition_penalty", "repetition_penalty", True),
    ("min_tokens", "min_tokens", True), ("seed", "seed", True), ("temp_scheduler", "temp_scheduler", True),
    ("extra_body", "extra_body", True), ("eval_config", "eval", False), ("val_config", "val", False),
    ("buffer_config", "buffer", False), ("learning_rate", "learning_rate", True), ("lora_alpha", "lora_alpha", True),
    ("oversampling_factor", "oversampling_factor", True), ("max_async_level", "max_async_level", True),
    ("checkpoints_config", "checkpoints", False), ("adapters_config", "adapters", False),
    ("checkpoint_id", "checkpoint_id", False), ("cluster_name", "cluster_name", False),
]  # fmt: skip


def build_run_payload(model_name: str, environments: list[dict[str, Any]], *, rollouts_per_example: int = 8,
                      max_steps: int = 100, batch_size: int = 128, secrets: dict[str, str] | None = None,
                      wandb_entity: str | None = None, wandb_project: str | None = None,
                      wandb_run_name: str | None = None, infrastructure_config: dict[str, Any] | None = None,
                      **optional: Any) -> dict[str, Any]:  # fmt: skip
    """Assemble the POST /rft/runs body; ``None`` options are omitted from the wire."""
    payload: dict[str, Any] = {
        "model": {"name": model_name},
        "environments": environments,
        "rollouts_per_example": rollouts_per_example,
        "max_steps": max_steps,
        "batch_size": batch_size,
        "secrets": [{"key": k, "value": v} for k, v in (secrets or {}).items()],
    }
    wandb = {k: v for k, v in (("entity", wandb_entity), ("project", wandb_project), ("name", wandb_run_name)) if v}
    if wandb:
        payload["monitoring"] = {"wandb": wandb}
    for kw, wire, keep_falsy in _OPTIONAL_PAYLOAD:
        v = optional.pop(kw, None)
        if v is None or (not keep_falsy and not v):
            continue
        payload[wire] = v
    if optional:
        raise TypeError(f"unknown run options: {sorted(optional)}")
    if infrastructure_config and "compute_size" in infrastructure_config:
        payload["compute_size"] = infrastructure_config["compute_size"]
    return payload


class RLClient:
    def __init__(self, client: Any) -> None:
        self.client = client

    @staticmethod
    def _team(team_id: str | None) -> dict | None:
        return {"team_id": team_id} if team_id else None

    def list_models(self, team_id: str | None = None) -> list[RLModel]:
        with wrap("list RL models"):
            return [RLModel.model_validate(m) for m in self.client.get("/rft/models", params=self._team(team_id)).get("models", [])]

    def list_runs(self, team_id: str | None = None) -> list[RLRun]:
        with wrap("list RL runs"):
            return [RLRun.model_validate(r) for r in self.client.get("/rft/runs", params=self._team(team_id)).get("runs", [])]

    def create_run(self, model_name: str, environments: list[dict[str, Any]], **kw: Any) -> RLRun:
        payload = build_run_payload(model_name, environments, **kw)
        try:
            with wrap("create RL run"):
                return RLRun.model_validate(self.client.post("/rft/runs", json=payload).get("run"))
        except ValidationError:
            raise  # 422 details are rendered field by field by the CLI

    def get_run(self, run_id: str) -> RLRun:
        with wrap("get RL run"):
            return RLRun.model_validate(self.client.get(f"/rft/runs/{run_id}").get("run"))

    def stop_run(self, run_id: str) -> RLRun:
        with wrap("stop RL run"):
            return RLRun.model_validate(self.client.request("PUT", f"/rft/runs/{run_id}/stop").get("run"))

    def restart_run(self, run_id: str) -> RLRun:
        """Server restarts a RUNNING run from its latest checkpoint."""
        with wrap("restart RL run"):
            return RLRun.model_validate(self.client.request("PUT", f"/rft/runs/{run_id}/restart").get("run"))

    def delete_run(self, run_id: str) -> None:
        with wrap("delete RL run"):
            self.client.delete(f"/rft/runs/{run_id}")

    def list_checkpoints(self, run_id: str, status_filter: str | None = None) -> list[RLCheckpoint]:
        with wrap("list checkpoints"):
            params = {"status_filter": status_filter} if status_filter else None
            data = self.client.get(f"/rft/runs/{run_id}/checkpoints", params=params)
            return [RLCheckpoint.model_validate(c) for c in data.get("checkpoints", [])]

    def get_logs(self, run_id: str, tail_lines: int = 1000) -> str:
        with wrap("get RL run logs"):
            return self.client.get(f"/rft/runs/{run_id}/logs", params={"tail_lines": tail_lines}).get("logs", "")

    def get_metrics(self, run_id: str, min_step: int | None = None, max_step: int | None = None,
                    limit: int | None = None) -> list[dict[str, Any]]:  # fmt: skip
        params = {k: v for k, v in (("min_step", min_step), ("max_step", max_step), ("limit", limit)) if v is not None}
        with wrap("get RL run metrics"):
            return self.client.get(f"/rft/runs/{run_id}/metrics", params=params or None).get("metrics", [])

    def get_rollouts(self, run_id: str, step: int, page: int = 1, limit: int = 100) -> dict[str, Any]:
        with wrap("get RL run rollouts"):
            r = self.client.get(f"/rft/runs/{run_id}/samples", params={"page": page, "limit": limit, "step": step})
            return {"run_id": r.get("run_id", run_id), "samples": r.get("samples", []), "total": r.get("total", 0),
                    "page": r.get("page", page), "limit": r.get("limit", limit), "total_pages": r.get("total_pages", 0)}  # fmt: skip

    def get_progress(self, run_id: str) -> dict[str, Any]:
        with wrap("get RL run progress"):
            r = self.client.get(f"/rft/runs/{run_id}/progress")
            return {"latest_step": r.get("latestStep"), "steps_with_samples": r.get("stepsWithSamples", []),
                    "steps_with_distributions": r.get("stepsWithDistributions", []), "last_updated_at": r.get("lastUpdatedAt")}  # fmt: skip

    def get_distributions(self, run_id: str, distribution_type: str | None = None, step: int | None = None) -> dict[str, Any]:
        params = {k: v for k, v in (("type", distribution_type), ("step", step)) if v is not None}
        with wrap("get RL run distributions"):
            r = self.client.get(f"/rft/runs/{run_id}/distributions", params=params)
            return {"bins": r.get("bins", []), "step": r.get("step")}

    def get_environment_status(self, owner: str, name: str) -> dict[str, Any]:
        with wrap(f"get status for {owner}/{name}"):
            return self.client.get(f"/environmentshub/{owner}/{name}/status").get("data") or {}
