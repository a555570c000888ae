"""Conformance of ``prime sandbox create``'s GPU / VM / image rules with the reference CLI: what is refused before any request
is made, and what the request carries otherwise (scenarios: packages/prime/tests/test_sandbox_cli.py:12-242; harness is ours)."""

from types import SimpleNamespace

import pytest
from typer.testing import CliRunner

from prime_b200.platform.commands import sandbox as sandbox_mod
from prime_b200.platform.main import app
from prime_b200.platform.utils.formatters import strip_ansi

runner = CliRunner()


@pytest.fixture
def created(monkeypatch):
    """Requests that reached SandboxClient.create (the list stays empty when the command refused first)."""
    seen = []
    monkeypatch.setenv("PRIME_API_KEY", "dummy")

    def create(self, request):
        seen.append(request)
        return SimpleNamespace(id=f"sbx-{len(seen)}")

    monkeypatch.setattr(sandbox_mod.SandboxClient, "create", create)
    return seen


def create(*argv):
    r = runner.invoke(app, ["sandbox", "create", *argv, "--yes"])
    return r.exit_code, strip_ansi(r.output)


def test_gpu_vm_with_an_image(created):
    code, out = create("team-1/gpu-runtime:v1", "--vm", "--gpu-count", "1", "--gpu-type", "H100_80GB")
    assert code == 0, out
    assert "Successfully created sandbox sbx-1" in out and "VM: Enabled" in out and "GPUs: H100_80GB x1" in out
    assert "Docker Image: team-1/gpu-runtime:v1" in out
    req = created[0]
    assert (req.docker_image, req.gpu_count, req.gpu_type, req.vm) == ("team-1/gpu-runtime:v1", 1, "H100_80GB", True)
    code, out = create("python:3.11-slim", "--vm", "--gpu-count", "1", "--gpu-type", "H100_80GB")
    assert code == 0 and created[1].docker_image == "python:3.11-slim" and created[1].vm is True


def test_vm_without_gpu(created):
    code, out = create("user-1/vm-image:latest", "--vm")
    assert code == 0 and "Successfully created sandbox" in out and created[0].vm is True and created[0].gpu_count == 0


@pytest.mark.parametrize("argv, complaint", [
    (("--gpu-count", "1", "--gpu-type", "RTX_PRO_6000"), "GPUs require VM sandboxes."),  # GPUs but no --vm (and no image)
    (("python:3.11-slim", "--gpu-count", "1", "--gpu-type", "H100_80GB"), "GPUs require VM sandboxes."),
    (("--gpu-count", "1"), "GPU type is required when requesting GPUs."),
    (("--gpu-type", "H100_80GB"), "GPU type provided without GPUs."),
    ((), "Docker image is required."),
])  # fmt: skip
def test_refused_before_any_request(created, argv, complaint):
    code, out = create(*argv)
    assert code == 1 and complaint in out and "Successfully created sandbox" not in out and not created
