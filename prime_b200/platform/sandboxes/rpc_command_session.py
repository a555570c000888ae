"""Connect-RPC ``command_session.CommandSession/Start`` helpers for VM sandboxes
(reference: packages/prime-sandboxes/src/prime_sandboxes/rpc_command_session.py:60-108)."""

from __future__ import annotations

from . import rpc_schema

try:  # connect-python is optional for container-only users
    from connectrpc.method import IdempotencyLevel, MethodInfo

    START_METHOD = MethodInfo(name="Start", service_name="command_session.CommandSession",
                              input=rpc_schema.StartRequest, output=rpc_schema.StartResponse,
                              idempotency_level=IdempotencyLevel.UNKNOWN)  # fmt: skip
except Exception:  # pragma: no cover
    START_METHOD = None


def build_start_request(command: str, working_dir: str | None, env: dict[str, str] | None):
    """Every command runs as ``/bin/bash -c <command>`` with stdin closed."""
    spec = rpc_schema.CommandSpec(cmd="/bin/bash", args=["-c", command], envs=env or {})
    if working_dir is not None:
        spec.cwd = working_dir
    return rpc_schema.StartRequest(command=spec, stdin=False)


class OutputCollector:
    """Accumulates a StartResponse event stream into stdout/stderr/exit code (PTY output counts as stdout)."""

    def __init__(self) -> None:
        self.stdout: list[str] = []
        self.stderr: list[str] = []
        self.exit_code: int | None = None

    def feed(self, response) -> None:
        if not response.HasField("event"):
            return
        ev = response.event
        kind = ev.WhichOneof("event")
        if kind == "end":
            self.exit_code = int(ev.end.exit_code)
        elif kind == "data":
            which = ev.data.WhichOneof("output")
            payload = getattr(ev.data, which, b"") if which else b""
            if payload:
                (self.stderr if which == "stderr" else self.stdout).append(payload.decode("utf-8", errors="replace"))

    def result(self) -> tuple[str, str, int | None]:
        return "".join(self.stdout), "".join(self.stderr), self.exit_code


# ---- the reference's functional spelling of the same two steps (rpc_command_session.py:69-108), for code written against it
COMMAND_SESSION_START_RPC_METHOD = START_METHOD
build_command_session_start_request = build_start_request


def collect_command_session_start_event(response, stdout_parts: list[str], stderr_parts: list[str]) -> int | None:
    """Append the event's output to the caller's lists; returns the exit code once the ``end`` event arrives, else ``None``."""
    c = OutputCollector()
    c.feed(response)
    stdout_parts.extend(c.stdout)
    stderr_parts.extend(c.stderr)
    return c.exit_code

