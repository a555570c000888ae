"""Evaluations SDK wire models (same fields as the reference: packages/prime-evals/src/prime_evals/models.py:8-135;
create request and stored record share one description base instead of repeating it)."""

from __future__ import annotations

from datetime import datetime
from enum import Enum
from typing import Any

from pydantic import BaseModel, ConfigDict, Field

from ..api._base import ApiModel


class EvaluationStatus(str, Enum):
    PENDING = "PENDING"
    RUNNING = "RUNNING"
    COMPLETED = "COMPLETED"
    FAILED = "FAILED"
    CANCELLED = "CANCELLED"


class _Described(BaseModel):
    """What an evaluation is *about* — the part a creator supplies and the API echoes back unchanged."""

    name: str
    model_name: str | None = None
    dataset: str | None = None
    framework: str | None = None
    task_type: str | None = None
    description: str | None = None
    tags: list[str] = Field(default_factory=list)
    metadata: dict[str, Any] | None = None
    metrics: dict[str, Any] | None = None
    suite_id: str | None = None
    run_id: str | None = None


class EnvironmentReference(BaseModel):
    id: str
    version_id: str | None = None


class CreateEvaluationRequest(_Described):
    """``POST /evaluations/``: the description plus what it ran on (environments, or a suite / training run)."""

    environments: list[dict[str, str]] | None = None


class Evaluation(_Described, ApiModel):
    """An evaluation as stored: the description plus identity, state and bookkeeping the server adds."""

    id: str = Field(..., alias="evaluation_id")
    status: str | None = None
    eval_type: str | None = None
    environment_ids: list[str] | None = None
    version_id: str | None = None
    total_samples: int | None = None
    created_at: datetime | None = None
    updated_at: datetime | None = None
    finalized_at: datetime | None = None
    user_id: str | None = None
    team_id: str | None = None


class Sample(ApiModel):
    example_id: int | None = None
    task: str | None = None
    prompt: list[dict[str, str]] | None = None
    completion: list[dict[str, str]] | None = None
    answer: str | None = None
    reward: float | None = None
    score: float | None = None
    correct: bool | None = None
    format_reward: float | None = None
    correctness: float | None = None
    info: dict[str, Any] | None = None


class SamplesResponse(ApiModel):
    samples: list[Sample]
    total: int
    page: int
    limit: int
    has_more: bool


class EvaluationListResponse(BaseModel):
    model_config = ConfigDict(populate_by_name=True)
    evaluations: list[Evaluation]
    total: int
    skip: int
    limit: int


class Environment(BaseModel):
    model_config = ConfigDict(populate_by_name=True, extra="allow")
    id: str
    name: str
    owner: str | None = None
    version: str | None = None
    description: str | None = None


class PushSamplesRequest(BaseModel):
    """Body of ``POST /evaluations/{id}/samples`` (reference: packages/prime-evals/src/prime_evals/models.py:114-117)."""

    samples: list[dict[str, Any]]


class FinalizeEvaluationRequest(BaseModel):
    """Body of ``POST /evaluations/{id}/finalize`` (reference: packages/prime-evals/src/prime_evals/models.py:120-123)."""

    metrics: dict[str, Any] | None = None

