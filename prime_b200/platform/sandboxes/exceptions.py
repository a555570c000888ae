"""Typed sandbox failures (reference: packages/prime-sandboxes/src/prime_sandboxes/exceptions.py:6-78)."""

from __future__ import annotations

from ..core.client import APIError


class SandboxFileNotFoundError(APIError):
    """404 from the gateway's read-file endpoint."""


class SandboxNotRunningError(RuntimeError):
    def __init__(self, sandbox_id: str, status: str | None = None, error_type: str | None = None,
                 command: str | None = None, message: str | None = None):  # fmt: skip
        self.sandbox_id, self.status, self.error_type, self.command = sandbox_id, status, error_type, command
        if not message:
            if error_type:
                message = f"Sandbox {sandbox_id} failed ({error_type})"
            elif status:
                message = f"Sandbox {sandbox_id} is not running (status={status})"
            else:
                message = f"Sandbox {sandbox_id} is not running"
        super().__init__(message)


class SandboxOOMError(SandboxNotRunningError):
    pass


class SandboxTimeoutError(SandboxNotRunningError):
    pass


class SandboxImagePullError(SandboxNotRunningError):
    pass


class _OpTimeout(RuntimeError):
    template = "{what} timed out after {timeout}s in sandbox {sandbox_id}"

    target_name = "target"  # the reference's name of the second constructor argument / attribute (command, file_path)

    def __init__(self, sandbox_id: str, target: str | None = None, timeout: int | None = None, **named: str):
        if target is None and self.target_name in named:
            target = named.pop(self.target_name)
        if named or target is None or timeout is None:
            raise TypeError(f"{type(self).__name__}(sandbox_id, {self.target_name}, timeout)")
        self.sandbox_id, self.target, self.timeout = sandbox_id, target, timeout
        setattr(self, self.target_name, target)
        super().__init__(self.template.format(what=self.describe(target), timeout=timeout, sandbox_id=sandbox_id))

    def describe(self, target: str) -> str:
        return target


class CommandTimeoutError(_OpTimeout):
    target_name = "command"

    def describe(self, target: str) -> str:
        return f"Command '{target}'"


class UploadTimeoutError(_OpTimeout):
    target_name = "file_path"

    def describe(self, target: str) -> str:
        return f"Upload to '{target}'"


class DownloadTimeoutError(_OpTimeout):
    target_name = "file_path"

    def describe(self, target: str) -> str:
        return f"Download from '{target}'"


ERROR_TYPE_TO_EXC = {"OOM_KILLED": SandboxOOMError, "TIMEOUT": SandboxTimeoutError, "IMAGE_PULL_FAILED": SandboxImagePullError}
