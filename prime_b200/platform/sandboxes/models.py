"""Sandbox SDK data models (reference: packages/prime-sandboxes/src/prime_sandboxes/models.py:10-263)."""

from __future__ import annotations

from datetime import datetime
from enum import Enum
from typing import Any

from pydantic import BaseModel, ConfigDict, Field, create_model, model_validator

from ..api._base import ApiModel


class SandboxStatus(str, Enum):
    PENDING = "PENDING"
    PROVISIONING = "PROVISIONING"
    RUNNING = "RUNNING"
    PAUSED = "PAUSED"
    ERROR = "ERROR"
    TERMINATED = "TERMINATED"
    TIMEOUT = "TIMEOUT"

    @classmethod
    def terminal(cls) -> tuple[str, ...]:
        return (cls.ERROR.value, cls.TERMINATED.value, cls.TIMEOUT.value)


class AdvancedConfigs(BaseModel):
    model_config = ConfigDict(extra="allow")


class _Wire(BaseModel):
    """Request bodies go out in snake_case, ``None`` omitted."""

    def wire(self) -> dict[str, Any]:
        return self.model_dump(by_alias=False, exclude_none=True)


class _Stamps(ApiModel):
    """Who owns a record and when it changed — shared by sandboxes and registry credentials."""

    created_at: datetime
    updated_at: datetime
    user_id: str | None = None
    team_id: str | None = None


class Sandbox(_Stamps):
    """One sandbox as the API returns it (camelCase on the wire; ``ApiModel`` supplies the aliases)."""

    # identity + state
    id: str
    name: str
    status: str
    labels: list[str] = Field(default_factory=list)
    # what runs
    docker_image: str
    start_command: str | None = None
    environment_vars: dict[str, Any] | None = None
    secrets: dict[str, Any] | None = None
    registry_credentials_id: str | None = None
    advanced_configs: AdvancedConfigs | None = None
    # what it runs on
    cpu_cores: float
    memory_gb: float = Field(..., alias="memoryGB")
    disk_size_gb: float = Field(..., alias="diskSizeGB")
    disk_mount_path: str
    gpu_count: int
    gpu_type: str | None = None
    vm: bool = False
    network_access: bool = True
    kubernetes_job_id: str | None = None
    # lifecycle
    timeout_minutes: int
    started_at: datetime | None = None
    terminated_at: datetime | None = None
    exit_code: int | None = None
    error_type: str | None = None
    error_message: str | None = None


class SandboxListResponse(ApiModel):
    sandboxes: list[Sandbox]
    total: int
    page: int
    per_page: int
    has_next: bool


class CreateSandboxRequest(_Wire):
    name: str
    docker_image: str
    start_command: str | None = "tail -f /dev/null"
    environment_vars: dict[str, str] | None = None
    secrets: dict[str, str] | None = None
    registry_credentials_id: str | None = None
    cpu_cores: float = 1.0
    memory_gb: float = 2.0
    disk_size_gb: float = 5.0
    gpu_count: int = 0
    gpu_type: str | None = None
    vm: bool = False
    network_access: bool = True
    timeout_minutes: int = 60
    labels: list[str] = Field(default_factory=list)
    team_id: str | None = None
    advanced_configs: AdvancedConfigs | None = None

    @model_validator(mode="after")
    def _gpu_rules(self) -> "CreateSandboxRequest":
        if self.gpu_count > 0:
            if not self.gpu_type:
                raise ValueError("gpu_type is required when gpu_count is greater than 0")
            if not self.vm:
                raise ValueError("gpu_count is only supported when vm is true")
        elif self.gpu_type is not None:
            raise ValueError("gpu_type requires gpu_count greater than 0")
        return self


# A PATCH body is the create body with everything optional, minus what cannot change after creation.
_FROZEN_AFTER_CREATE = ("vm", "labels", "team_id", "advanced_configs")
UpdateSandboxRequest = create_model(  # type: ignore[call-overload]
    "UpdateSandboxRequest",
    __base__=_Wire,
    __module__=__name__,
    **{
        field: (info.annotation | None, None)
        for field, info in CreateSandboxRequest.model_fields.items()
        if field not in _FROZEN_AFTER_CREATE
    },
)


class CommandRequest(_Wire):
    command: str
    working_dir: str | None = None
    env: dict[str, str] | None = None


class CommandResponse(BaseModel):
    stdout: str
    stderr: str
    exit_code: int


class FileUploadResponse(BaseModel):
    success: bool
    path: str
    size: int
    timestamp: datetime


class ReadFileResponse(BaseModel):
    content: str
    size: int


class SandboxLogsResponse(BaseModel):
    logs: str


class BulkDeleteSandboxRequest(_Wire):
    sandbox_ids: list[str] | None = None
    labels: list[str] | None = None


class BulkDeleteSandboxResponse(BaseModel):
    succeeded: list[str]
    failed: list[dict[str, str]]
    message: str


class RegistryCredentialSummary(_Stamps):
    id: str
    name: str
    server: str


class DockerImageCheckResponse(BaseModel):
    accessible: bool
    details: str


class ExposePortRequest(_Wire):
    port: int
    name: str | None = None
    protocol: str = "HTTP"  # HTTP | TCP


class ExposedPort(BaseModel):
    exposure_id: str
    sandbox_id: str
    port: int
    name: str | None = None
    url: str
    tls_socket: str
    protocol: str | None = None
    external_port: int | None = None
    external_endpoint: str | None = None
    created_at: str | None = None


class ListExposedPortsResponse(BaseModel):
    exposures: list[ExposedPort]


class SSHSession(BaseModel):
    session_id: str
    exposure_id: str
    sandbox_id: str
    host: str
    port: int
    external_endpoint: str
    expires_at: datetime
    ttl_seconds: int
    gateway_url: str
    user_ns: str
    job_id: str
    token: str


class BackgroundJob(BaseModel):
    job_id: str
    sandbox_id: str
    stdout_log_file: str
    stderr_log_file: str
    exit_file: str


class BackgroundJobStatus(BaseModel):
    job_id: str
    completed: bool
    exit_code: int | None = None
    stdout: str | None = None
    stderr: str | None = None
