"""Conformance of the sandbox SDK's wire models with the reference — defaults, the GPU ⇒ VM + type rules, enum values, and
parsing the API's camelCase answer; plus the derived PATCH model (ours) staying in step with the create model
(scenarios: packages/prime-sandboxes/tests/test_models.py:13-130; harness is ours)."""

import pytest
from pydantic import ValidationError

from prime_b200.platform.sandboxes import CreateSandboxRequest, Sandbox, SandboxStatus, UpdateSandboxRequest


def test_create_defaults():
    r = CreateSandboxRequest(name="test-sandbox", docker_image="python:3.11-slim")
    assert (r.cpu_cores, r.memory_gb, r.disk_size_gb, r.gpu_count, r.gpu_type, r.vm, r.timeout_minutes, r.labels) == (1, 2, 5, 0, None, False, 60, [])
    assert CreateSandboxRequest(name="x", docker_image="y", gpu_type=None).gpu_type is None  # an explicit None is the default


@pytest.mark.parametrize("kw", [{"gpu_count": 1, "vm": True}, {"gpu_count": 1, "gpu_type": "H100_80GB"}, {"gpu_type": "H100_80GB"}])
def test_gpu_rules_reject(kw):
    with pytest.raises(ValidationError):
        CreateSandboxRequest(name="gpu", docker_image="img", **kw)


def test_gpu_rules_accept():
    r = CreateSandboxRequest(name="gpu", docker_image="img", gpu_count=1, gpu_type="H100_80GB", vm=True)
    assert (r.gpu_count, r.gpu_type, r.vm) == (1, "H100_80GB", True)


def test_status_values_are_plain_strings():
    assert SandboxStatus.PENDING == "PENDING" and SandboxStatus.RUNNING == "RUNNING" and SandboxStatus.TERMINATED == "TERMINATED"


def test_api_answer_parses_from_camel_case():
    s = Sandbox.model_validate({
        "id": "test-123", "name": "test-sandbox", "dockerImage": "python:3.11-slim", "startCommand": None, "cpuCores": 2, "memoryGB": 4,
        "diskSizeGB": 10, "diskMountPath": "/workspace", "gpuCount": 1, "gpuType": "H100_80GB", "vm": True, "status": "RUNNING",
        "timeoutMinutes": 60, "createdAt": "2025-01-01T00:00:00Z", "updatedAt": "2025-01-01T00:00:00Z"})  # fmt: skip
    assert (s.id, s.name, s.cpu_cores, s.memory_gb, s.status, s.gpu_type, s.vm) == ("test-123", "test-sandbox", 2, 4, "RUNNING", "H100_80GB", True)
    assert s.network_access is True and s.labels == [] and s.team_id is None  # what the answer may leave out


def test_patch_model_is_the_create_model_made_optional():
    fields = set(UpdateSandboxRequest.model_fields)
    assert fields == set(CreateSandboxRequest.model_fields) - {"vm", "labels", "team_id", "advanced_configs"}
    assert UpdateSandboxRequest().wire() == {} and UpdateSandboxRequest(cpu_cores=2, gpu_count=1).wire() == {"cpu_cores": 2.0, "gpu_count": 1}
