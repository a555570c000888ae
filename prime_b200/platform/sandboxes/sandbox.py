"""Sandbox SDK: sync + async clients for remote code-execution sandboxes.

Parity map (reference packages/prime-sandboxes/src/prime_sandboxes/sandbox.py):
  retry policy split by idempotency ............ :66-109   → ``GatewayPolicy``
  terminated / OOM / timeout classification ..... :127-200  → ``classify_not_running``
  auth-token cache, expiry pruning, single-flight :203-431  → ``AuthCache`` / ``AsyncAuthCache``
  SandboxClient ................................. :458-1272 → ``SandboxClient``
  AsyncSandboxClient (shared pooled AsyncClient) . :1275-2153 → ``AsyncSandboxClient``
  TemplateClient ................................ :2156-2215 → ``TemplateClient`` / ``AsyncTemplateClient``

Design difference: everything that is *not* I/O (URL building, payloads, error → exception mapping, the
409 decision, background-job shell wrapping, status aggregation) lives once in module-level pure
functions; the sync and async classes only differ in how they perform the request.
"""

from __future__ import annotations

import asyncio
import functools
import json
import os
import re
import shlex
import threading
import time
import uuid
from datetime import datetime, timezone
from pathlib import Path
from typing import Any, NoReturn

import httpx

from ..core.client import TRANSPORT_RETRY, APIClient, APIError, AsyncAPIClient, user_agent
from .exceptions import (
    ERROR_TYPE_TO_EXC,
    CommandTimeoutError,
    DownloadTimeoutError,
    SandboxFileNotFoundError,
    SandboxNotRunningError,
    UploadTimeoutError,
)
from .models import (
    BackgroundJob,
    BackgroundJobStatus,
    BulkDeleteSandboxRequest,
    BulkDeleteSandboxResponse,
    CommandResponse,
    CreateSandboxRequest,
    DockerImageCheckResponse,
    ExposedPort,
    ExposePortRequest,
    FileUploadResponse,
    ListExposedPortsResponse,
    ReadFileResponse,
    RegistryCredentialSummary,
    Sandbox,
    SandboxListResponse,
    SandboxLogsResponse,
    SandboxStatus,
    SSHSession,
)

# --------------------------------------------------------------------------------------- policy
CONNECT_ERRORS = (httpx.RemoteProtocolError, httpx.ConnectError, httpx.PoolTimeout)  # NOT ReadTimeout: may have executed
RETRYABLE_5XX = frozenset({500, 502, 503, 504, 524})
MAX_409_RETRIES = 4
RETRY_409_BASE_DELAY = 0.25
GATEWAY_ATTEMPTS = 4
DEFAULT_COMMAND_TIMEOUT = 300
_ENV_KEY = re.compile(r"[A-Za-z_][A-Za-z0-9_]*\Z")
NOT_FOUND_HINT = "Sandbox is no longer present on the runtime node. Please create a new sandbox."


def __getattr__(name: str):
    """``sandbox.ConnectClientSync`` / ``sandbox.ConnectClient``: the connectrpc client classes, imported on first use (only VM sandboxes pay
    for connectrpc + protobuf) yet still module attributes — code that substitutes a fake transport there keeps working."""
    if name in ("ConnectClientSync", "ConnectClient"):
        import connectrpc.client

        return getattr(connectrpc.client, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def _rpc_class(name: str):
    return globals().get(name) or __getattr__(name)


def sandboxes_user_agent() -> str:
    return user_agent("prime-b200-sandboxes")


def gateway_says_sandbox_gone(resp: httpx.Response | None) -> bool:
    if resp is None or resp.status_code != 502:
        return False
    try:
        body = resp.json()
    except Exception:
        return False
    return isinstance(body, dict) and body.get("error") == "sandbox_not_found"


def retryable(exc: BaseException, idempotent: bool) -> bool:
    """Idempotent requests also retry 5xx (unless the gateway says the sandbox is gone); POSTs only
    retry failures that prove the request never reached the server."""
    if isinstance(exc, CONNECT_ERRORS):
        return True
    if idempotent and isinstance(exc, httpx.HTTPStatusError) and exc.response.status_code in RETRYABLE_5XX:
        return not gateway_says_sandbox_gone(exc.response)
    return False


def backoff(attempt: int, lo: float = 1.0, hi: float = 30.0) -> float:
    return max(lo, min(hi, float(2**attempt)))


def gateway_retry(idempotent: bool, sleep=None):
    """Decorator form of the gateway policy: up to ``GATEWAY_ATTEMPTS`` calls, 1 → 30 s exponential back-off between them, retrying only what
    ``retryable(exc, idempotent)`` allows; the last failure propagates unchanged.  ``sleep`` defaults to ``time.sleep`` looked up at call time."""

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*a, **kw):
            for attempt in range(GATEWAY_ATTEMPTS):
                try:
                    return fn(*a, **kw)
                except Exception as e:
                    if attempt + 1 < GATEWAY_ATTEMPTS and retryable(e, idempotent):
                        (sleep or time.sleep)(backoff(attempt))
                        continue
                    raise
            raise AssertionError("unreachable")

        return wrapper

    return deco


# the reference module's names for the two policies and their inputs (prime_sandboxes/sandbox.py:66-109)
GATEWAY_RETRYABLE_EXCEPTIONS = CONNECT_ERRORS
RETRYABLE_5XX_STATUSES = RETRYABLE_5XX
_gateway_retry = gateway_retry(idempotent=True)
_gateway_post_retry = gateway_retry(idempotent=False)


def validate_env_key(key: str) -> str:
    if not _ENV_KEY.match(key):
        raise ValueError(f"Invalid environment variable name: {key!r}")
    return key


# --------------------------------------------------------------------------------------- failure classification
def describe_dead_sandbox(command: str, ctx: dict) -> str:
    preview = command if len(command) <= 50 else command[:50] + "..."
    parts = [f"Command '{preview}' failed: sandbox is no longer running."]
    et = ctx.get("error_type")
    if et == "OOM_KILLED":
        parts += ["The sandbox was terminated due to out-of-memory (OOM).",
                  "Consider requesting more memory or optimizing memory usage."]  # fmt: skip
    elif et == "TIMEOUT":
        parts.append("The sandbox exceeded its maximum runtime and was terminated.")
    elif et == "IMAGE_PULL_FAILED":
        parts.append("The sandbox failed to start due to image pull failure.")
    elif ctx.get("status") == "TERMINATED":
        parts.append("The sandbox was terminated.")
    if ctx.get("error_message"):
        parts.append(f"Details: {ctx['error_message']}")
    return " ".join(parts)


def classify_not_running(sandbox_id: str, ctx: dict, command: str | None = None,
                         cause: BaseException | None = None) -> NoReturn:  # fmt: skip
    et, status = ctx.get("error_type"), ctx.get("status")
    if command:
        message = describe_dead_sandbox(command, ctx)
    elif ctx.get("error_message"):
        message = f"Sandbox {sandbox_id} failed ({et}): {ctx['error_message']}"
    else:
        message = None
    exc = ERROR_TYPE_TO_EXC.get(et, SandboxNotRunningError)(sandbox_id, status, et, command=command, message=message)
    if cause is not None:
        raise exc from cause
    raise exc


def mark_gone(ctx: dict) -> dict:
    ctx = dict(ctx)
    ctx["status"] = "TERMINATED"
    ctx["error_type"] = ctx.get("error_type") or "SANDBOX_NOT_FOUND"
    ctx["error_message"] = ctx.get("error_message") or NOT_FOUND_HINT
    return ctx


def parse_error_context(resp: dict | None) -> dict:
    resp = resp or {}
    return {"status": resp.get("status"), "error_type": resp.get("errorType") or resp.get("error_type"),
            "error_message": resp.get("errorMessage") or resp.get("error_message")}  # fmt: skip


EMPTY_CTX = {"status": None, "error_type": None, "error_message": None}


def http_detail(e: httpx.HTTPStatusError) -> str:
    return f"HTTP {e.response.status_code} {e.request.method} {e.request.url}: {e.response.text}"


def request_detail(e: httpx.RequestError) -> str:
    req = getattr(e, "_request", None)
    where = f"{req.method} {req.url}" if req is not None else "?"
    return f"{type(e).__name__} at {where}: {e}"


# --------------------------------------------------------------------------------------- pure builders
def gateway_url(auth: dict, suffix: str = "") -> str:
    base = f"{auth['gateway_url'].rstrip('/')}/{auth['user_ns']}/{auth['job_id']}"
    return f"{base}/{suffix}" if suffix else base


def bearer(auth: dict) -> dict[str, str]:
    return {"Authorization": f"Bearer {auth['token']}"}


def exec_payload(sandbox_id: str, command: str, working_dir: str | None, env: dict | None, timeout: int) -> dict:
    return {"command": command, "working_dir": working_dir, "env": env or {}, "sandbox_id": sandbox_id, "timeout": timeout}


def background_command(command: str, working_dir: str | None, env: dict[str, str] | None) -> tuple[BackgroundJob, str]:
    """nohup wrapper: output to log files, exit code to a marker file.  The user command runs in a subshell
    so an ``exit`` inside it cannot skip the ``echo $?``."""
    job_id = uuid.uuid4().hex[:8]
    out, err, code = (f"/tmp/job_{job_id}.{s}" for s in ("stdout.log", "stderr.log", "exit"))
    exports = "".join(f"export {validate_env_key(k)}={shlex.quote(v)}; " for k, v in (env or {}).items())
    cd = f"cd {shlex.quote(working_dir)} && " if working_dir else ""
    inner = f"({exports}{cd}{command}) > {shlex.quote(out)} 2> {shlex.quote(err)}; echo $? > {shlex.quote(code)}"
    shell = f"nohup sh -c {shlex.quote(inner)} < /dev/null > /dev/null 2>&1 &"
    return BackgroundJob(job_id=job_id, sandbox_id="", stdout_log_file=out, stderr_log_file=err, exit_file=code), shell


def parse_exit_marker(content: str) -> int | None:
    s = content.strip()
    try:
        return int(s) if s else None
    except ValueError:
        return None


def tally_statuses(sandboxes: list[Sandbox], wanted: set[str]) -> tuple[int, list[tuple[str, str]], dict[str, str]]:
    running, failed, statuses = 0, [], {}
    for s in sandboxes:
        if s.id not in wanted:
            continue
        statuses[s.id] = s.status
        if s.status == SandboxStatus.RUNNING.value:
            running += 1
        elif s.status in SandboxStatus.terminal():
            failed.append((s.id, s.status))
    return running, failed, statuses


def is_rate_limited(e: Exception) -> bool:
    return "429" in str(e) or "Too Many Requests" in str(e)


# --------------------------------------------------------------------------------------- auth cache
def _expiry(info: dict) -> datetime:
    dt = datetime.fromisoformat(info["expires_at"].replace("Z", "+00:00"))
    return dt if dt.tzinfo else dt.replace(tzinfo=timezone.utc)


def _still_valid(info: dict) -> bool:
    try:
        return datetime.now(timezone.utc) < _expiry(info)
    except Exception:
        return False


class _CacheStore:
    """On-disk token cache (``~/.prime/sandbox_auth_cache.json``); expired entries are pruned on load."""

    def __init__(self, path: Path):
        self.path = Path(path)
        self.entries: dict[str, dict] = {}
        try:
            raw = json.loads(self.path.read_text()) if self.path.exists() else {}
        except Exception:
            raw = {}
        self.entries = {k: v for k, v in raw.items() if isinstance(v, dict) and _still_valid(v)}
        if len(self.entries) != len(raw):
            self.save()

    def save(self) -> None:
        try:
            self.path.parent.mkdir(parents=True, exist_ok=True)
            tmp = self.path.with_suffix(f".{os.getpid()}.tmp")
            tmp.write_text(json.dumps(self.entries))
            os.replace(tmp, self.path)
        except Exception:
            pass

    def fresh(self, sandbox_id: str) -> dict | None:
        info = self.entries.get(sandbox_id)
        if info is None:
            return None
        if _still_valid(info):
            return dict(info)
        del self.entries[sandbox_id]
        return None


class AuthCache:
    """Thread-safe; concurrent misses for one sandbox coalesce into a single POST /sandbox/{id}/auth."""

    def __init__(self, path: Path, client: APIClient):
        self.client = client
        self._lock = threading.Lock()
        self._inflight: dict[str, threading.Event] = {}
        self._store = _CacheStore(path)

    def get_or_refresh(self, sandbox_id: str) -> dict:
        while True:
            with self._lock:
                hit = self._store.fresh(sandbox_id)
                if hit:
                    return hit
                waiter = self._inflight.get(sandbox_id)
                if waiter is None:
                    self._inflight[sandbox_id] = threading.Event()
            if waiter is not None:  # someone else is fetching: wait, then re-check
                waiter.wait()
                continue
            try:
                fresh = self.client.request("POST", f"/sandbox/{sandbox_id}/auth")
                with self._lock:
                    self._store.entries[sandbox_id] = fresh
                    self._store.save()
                return dict(fresh)
            finally:
                with self._lock:
                    ev = self._inflight.pop(sandbox_id, None)
                if ev:
                    ev.set()

    def is_vm(self, sandbox_id: str) -> bool:
        with self._lock:
            hit = self._store.fresh(sandbox_id)
        if hit and isinstance(hit.get("is_vm"), bool):
            return hit["is_vm"]
        vm = Sandbox.model_validate(self.client.request("GET", f"/sandbox/{sandbox_id}")).vm
        with self._lock:
            if sandbox_id in self._store.entries:
                self._store.entries[sandbox_id]["is_vm"] = vm
                self._store.save()
        return vm

    def set(self, sandbox_id: str, info: dict) -> None:
        with self._lock:
            self._store.entries[sandbox_id] = info
            self._store.save()

    def clear(self) -> None:
        with self._lock:
            self._store.entries = {}
            self._store.save()


class AsyncAuthCache:
    def __init__(self, path: Path, client: AsyncAPIClient):
        self.client = client
        self._lock = asyncio.Lock()
        self._inflight: dict[str, asyncio.Event] = {}
        self._store = _CacheStore(path)

    async def get_or_refresh(self, sandbox_id: str) -> dict:
        while True:
            async with self._lock:
                hit = self._store.fresh(sandbox_id)
                if hit:
                    return hit
                waiter = self._inflight.get(sandbox_id)
                if waiter is None:
                    self._inflight[sandbox_id] = asyncio.Event()
            if waiter is not None:
                await waiter.wait()
                continue
            try:
                fresh = await self.client.request("POST", f"/sandbox/{sandbox_id}/auth")
                async with self._lock:
                    self._store.entries[sandbox_id] = fresh
                    self._store.save()
                return dict(fresh)
            finally:
                async with self._lock:
                    ev = self._inflight.pop(sandbox_id, None)
                if ev:
                    ev.set()

    async def is_vm(self, sandbox_id: str) -> bool:
        async with self._lock:
            hit = self._store.fresh(sandbox_id)
        if hit and isinstance(hit.get("is_vm"), bool):
            return hit["is_vm"]
        vm = Sandbox.model_validate(await self.client.request("GET", f"/sandbox/{sandbox_id}")).vm
        async with self._lock:
            if sandbox_id in self._store.entries:
                self._store.entries[sandbox_id]["is_vm"] = vm
                self._store.save()
        return vm

    async def set(self, sandbox_id: str, info: dict) -> None:
        async with self._lock:
            self._store.entries[sandbox_id] = info
            self._store.save()

    async def clear(self) -> None:
        async with self._lock:
            self._store.entries = {}
            self._store.save()


# the reference's public names of the two caches (sandbox.py: SandboxAuthCache / AsyncSandboxAuthCache)
SandboxAuthCache = AuthCache
AsyncSandboxAuthCache = AsyncAuthCache


# --------------------------------------------------------------------------------------- sync client
class SandboxClient:
    """All gateway traffic of one client shares ONE pooled, thread-safe ``httpx.Client`` with per-request timeouts (keep-alive
    connections, one TLS context). The reference builds a fresh ``httpx.Client`` for every gateway request — a new TLS context
    (≈20 ms of ``load_verify_locations``) and a new connection per command; `tools/platform_bench.py` measures the difference."""

    def __init__(self, api_client: APIClient | None = None, max_connections: int = 100, max_keepalive_connections: int = 32):
        self.client = api_client or APIClient(user_agent=sandboxes_user_agent(), retry=TRANSPORT_RETRY)
        self._auth_cache = AuthCache(self.client.config.config_dir / "sandbox_auth_cache.json", self.client)
        self._sleep = time.sleep  # injectable for tests
        self._limits = httpx.Limits(max_connections=max_connections, max_keepalive_connections=max_keepalive_connections)
        self._gw: httpx.Client | None = None
        self._gw_lock = threading.Lock()

    def _pool(self) -> httpx.Client:
        gw = self._gw
        if gw is None or gw.is_closed:
            with self._gw_lock:
                gw = self._gw
                if gw is None or gw.is_closed:
                    gw = self._gw = httpx.Client(limits=self._limits, timeout=None)
        return gw

    def close(self) -> None:
        """Release the pooled gateway connections (the client stays usable: the pool is rebuilt on demand)."""
        with self._gw_lock:
            gw, self._gw = self._gw, None
        if gw is not None and not gw.is_closed:
            gw.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- gateway I/O with the idempotency-aware retry split
    def _gateway(self, method: str, url: str, *, idempotent: bool, headers: dict, timeout: float, **kw) -> httpx.Response:
        @gateway_retry(idempotent, sleep=self._sleep)
        def once() -> httpx.Response:
            resp = self._pool().request(method, url, headers=headers, timeout=timeout, **kw)
            if idempotent and resp.status_code in RETRYABLE_5XX:
                resp.raise_for_status()
            return resp

        return once()

    def _gateway_post(self, url: str, headers: dict, timeout: float, **kw) -> httpx.Response:
        return self._gateway("POST", url, idempotent=False, headers=headers, timeout=timeout, **kw)

    def _gateway_get(self, url: str, headers: dict, params: dict, timeout: float) -> httpx.Response:
        return self._gateway("GET", url, idempotent=True, headers=headers, timeout=timeout, params=params)

    # ---- failure helpers
    def _get_sandbox_error_context(self, sandbox_id: str) -> dict:
        try:
            return parse_error_context(self.client.request("GET", f"/sandbox/{sandbox_id}/error-context"))
        except Exception:
            return dict(EMPTY_CTX)

    def _should_retry_409(self, sandbox_id: str, error: httpx.HTTPStatusError, attempt: int, command: str | None = None) -> bool:
        """409 from the gateway: transient only if the control plane still reports RUNNING."""
        ctx = self._get_sandbox_error_context(sandbox_id)
        if ctx["status"] != "RUNNING":
            classify_not_running(sandbox_id, ctx, command=command, cause=error)
        if attempt < MAX_409_RETRIES - 1:
            self._sleep(RETRY_409_BASE_DELAY * (2**attempt))
            return True
        raise APIError(f"Sandbox {sandbox_id} returned 409 after {MAX_409_RETRIES} retries. "
                       "This may be a transient DNS or gateway issue. Please retry.") from error  # fmt: skip

    def _timeout_or_dead(self, sandbox_id: str, command: str, timeout: int, cause: BaseException) -> NoReturn:
        ctx = self._get_sandbox_error_context(sandbox_id)
        if ctx["status"] in SandboxStatus.terminal():
            classify_not_running(sandbox_id, ctx, command=command, cause=cause)
        raise CommandTimeoutError(sandbox_id, command, timeout) from cause

    def _is_sandbox_reachable(self, sandbox_id: str, timeout: int = 10) -> bool:
        try:
            self.execute_command(sandbox_id, "echo 'sandbox ready'", timeout=timeout)
            return True
        except Exception:
            return False

    def clear_auth_cache(self) -> None:
        self._auth_cache.clear()

    # ---- control plane CRUD
    def create(self, request: CreateSandboxRequest) -> Sandbox:
        if request.team_id is None:
            request.team_id = self.client.config.team_id
        return Sandbox.model_validate(self.client.request("POST", "/sandbox", json=request.wire()))

    def list(self, team_id: str | None = None, status: str | None = None, labels: list[str] | None = None, page: int = 1,
             per_page: int = 50, exclude_terminated: bool | None = None) -> SandboxListResponse:  # fmt: skip
        return SandboxListResponse.model_validate(
            self.client.request("GET", "/sandbox", params=_list_params(self.client, team_id, status, labels, page, per_page, exclude_terminated))
        )

    def get(self, sandbox_id: str) -> Sandbox:
        return Sandbox.model_validate(self.client.request("GET", f"/sandbox/{sandbox_id}"))

    def delete(self, sandbox_id: str) -> dict[str, Any]:
        return self.client.request("DELETE", f"/sandbox/{sandbox_id}")

    def bulk_delete(self, sandbox_ids: list[str] | None = None, labels: list[str] | None = None) -> BulkDeleteSandboxResponse:
        body = BulkDeleteSandboxRequest(sandbox_ids=sandbox_ids, labels=labels).wire()
        return BulkDeleteSandboxResponse.model_validate(self.client.request("DELETE", "/sandbox", json=body))

    def get_logs(self, sandbox_id: str) -> str:
        return SandboxLogsResponse.model_validate(self.client.request("GET", f"/sandbox/{sandbox_id}/logs")).logs

    # ---- command execution: VM → Connect-RPC stream, container → REST
    def execute_command(self, sandbox_id: str, command: str, working_dir: str | None = None, env: dict[str, str] | None = None,
                        timeout: int | None = None) -> CommandResponse:  # fmt: skip
        auth = self._auth_cache.get_or_refresh(sandbox_id)
        run = self._execute_command_connect_rpc if self._auth_cache.is_vm(sandbox_id) else self._execute_command_rest
        return run(sandbox_id=sandbox_id, command=command, auth=auth, working_dir=working_dir, env=env, timeout=timeout)

    def _execute_command_connect_rpc(self, sandbox_id: str, command: str, auth: dict, working_dir: str | None = None,
                                     env: dict[str, str] | None = None, timeout: int | None = None) -> CommandResponse:  # fmt: skip
        from connectrpc.code import Code
        from connectrpc.errors import ConnectError

        from .rpc_command_session import START_METHOD, OutputCollector, build_start_request  # protobuf: only VM sandboxes pay for the import

        limit = timeout if timeout is not None else DEFAULT_COMMAND_TIMEOUT
        rpc = _rpc_class("ConnectClientSync")(gateway_url(auth))
        out = OutputCollector()
        try:
            for ev in rpc.execute_server_stream(request=build_start_request(command, working_dir, env), method=START_METHOD,
                                                headers=bearer(auth), timeout_ms=limit * 1000):  # fmt: skip
                out.feed(ev)
            stdout, stderr, code = out.result()
            if code is None:
                raise APIError("Command stream ended without exit code")
            return CommandResponse(stdout=stdout, stderr=stderr, exit_code=code)
        except ConnectError as e:
            if e.code == Code.DEADLINE_EXCEEDED:
                self._timeout_or_dead(sandbox_id, command, limit, e)
            if e.code == Code.NOT_FOUND:
                classify_not_running(sandbox_id, mark_gone(self._get_sandbox_error_context(sandbox_id)), command=command, cause=e)
            raise APIError(f"Connect RPC failed ({e.code.value}): {e.message}") from e
        except (APIError, SandboxNotRunningError, CommandTimeoutError):
            raise
        except Exception as e:
            raise APIError(f"Request failed: {type(e).__name__}: {e}") from e
        finally:
            rpc.close()

    def _execute_command_rest(self, sandbox_id: str, command: str, auth: dict, working_dir: str | None = None,
                              env: dict[str, str] | None = None, timeout: int | None = None) -> CommandResponse:  # fmt: skip
        limit = timeout if timeout is not None else DEFAULT_COMMAND_TIMEOUT
        url, payload = gateway_url(auth, "exec"), exec_payload(sandbox_id, command, working_dir, env, limit)
        for attempt in range(MAX_409_RETRIES):
            try:
                # +5 s of slack so a command finishing right at its limit is not cut by the client first
                resp = self._gateway_post(url, headers=bearer(auth), timeout=limit + 5, json=payload)
                resp.raise_for_status()
                return CommandResponse.model_validate(resp.json())
            except httpx.TimeoutException as e:
                self._timeout_or_dead(sandbox_id, command, limit, e)
            except httpx.HTTPStatusError as e:
                code = e.response.status_code
                if gateway_says_sandbox_gone(e.response):
                    classify_not_running(sandbox_id, mark_gone(self._get_sandbox_error_context(sandbox_id)), command=command, cause=e)
                if code == 409 and self._should_retry_409(sandbox_id, e, attempt, command=command):
                    continue
                if code == 408:
                    self._timeout_or_dead(sandbox_id, command, limit, e)
                raise APIError(http_detail(e)) from e
            except httpx.RequestError as e:
                raise APIError(f"Request failed: {request_detail(e)}") from e
        raise APIError("Command execution failed after retries")

    # ---- background jobs (for commands longer than an HTTP request should live)
    def start_background_job(self, sandbox_id: str, command: str, working_dir: str | None = None,
                             env: dict[str, str] | None = None) -> BackgroundJob:  # fmt: skip
        job, shell = background_command(command, working_dir, env)
        job.sandbox_id = sandbox_id
        self.execute_command(sandbox_id, shell, timeout=10)
        return job

    def _read_or_empty(self, sandbox_id: str, path: str) -> str:
        try:
            return self.read_file(sandbox_id, path).content
        except SandboxFileNotFoundError:
            return ""

    def get_background_job(self, sandbox_id: str, job: BackgroundJob) -> BackgroundJobStatus:
        code = parse_exit_marker(self._read_or_empty(sandbox_id, job.exit_file))
        if code is None:
            return BackgroundJobStatus(job_id=job.job_id, completed=False)
        return BackgroundJobStatus(job_id=job.job_id, completed=True, exit_code=code,
                                   stdout=self._read_or_empty(sandbox_id, job.stdout_log_file),
                                   stderr=self._read_or_empty(sandbox_id, job.stderr_log_file))  # fmt: skip

    def run_background_job(self, sandbox_id: str, command: str, timeout: int = 900, working_dir: str | None = None,
                           env: dict[str, str] | None = None, poll_interval: int = 3) -> BackgroundJobStatus:  # fmt: skip
        job = self.start_background_job(sandbox_id, command, working_dir=working_dir, env=env)
        deadline = time.monotonic() + timeout
        while time.monotonic() < deadline:
            st = self.get_background_job(sandbox_id, job)
            if st.completed:
                return st
            self._sleep(poll_interval)
        raise CommandTimeoutError(sandbox_id, command, timeout)

    # ---- readiness
    def wait_for_creation(self, sandbox_id: str, max_attempts: int = 60, stability_checks: int = 1) -> None:
        streak = 0
        for attempt in range(max_attempts):
            sb = self.get(sandbox_id)
            if sb.status == "RUNNING":
                if self._is_sandbox_reachable(sandbox_id):
                    streak += 1
                    if streak >= stability_checks:
                        return
                    self._sleep(0.5)
                    continue
                streak = 0
            elif sb.status in SandboxStatus.terminal():
                classify_not_running(sb.id, {"status": sb.status, "error_type": sb.error_type, "error_message": sb.error_message})
            self._sleep(1 if attempt < 5 else 2)  # fast polls first, then back off
        raise SandboxNotRunningError(sandbox_id, "Timeout during sandbox creation")

    def bulk_wait_for_creation(self, sandbox_ids: list[str], max_attempts: int = 60) -> dict[str, str]:
        """Polls the LIST endpoint (one request per page) instead of N GETs, so large fleets stay under rate limits."""
        wanted, final = set(sandbox_ids), {}
        for attempt in range(max_attempts):
            running, failed, page = 0, [], 1
            while True:
                try:
                    resp = self.list(per_page=100, page=page)
                except Exception as e:
                    if is_rate_limited(e):
                        self._sleep(min(2**attempt, 60))
                        continue
                    raise
                r, f, st = tally_statuses(resp.sandboxes, wanted)
                running, failed = running + r, failed + f
                final.update(st)
                if len(final) == len(sandbox_ids) or not resp.has_next:
                    break
                page += 1
            if failed:
                raise RuntimeError(f"Sandboxes failed: {failed}")
            if running == len(sandbox_ids):
                unreachable = [s for s in sandbox_ids if final.get(s) == "RUNNING" and not self._is_sandbox_reachable(s)]
                for s in unreachable:
                    final.pop(s, None)
                if not unreachable:
                    return final
            self._sleep(1 if attempt < 5 else 2)
        for s in wanted - set(final):
            final[s] = "TIMEOUT"
        raise RuntimeError(f"Timeout waiting for sandboxes to be ready. Status: {final}")

    # ---- files
    def _upload(self, sandbox_id: str, file_path: str, filename: str, content: bytes, timeout: int | None) -> FileUploadResponse:
        auth = self._auth_cache.get_or_refresh(sandbox_id)
        limit = timeout if timeout is not None else 300
        for attempt in range(MAX_409_RETRIES):
            try:
                resp = self._gateway_post(gateway_url(auth, "upload"), headers=bearer(auth), timeout=limit,
                                          files={"file": (filename, content)}, params={"path": file_path, "sandbox_id": sandbox_id})  # fmt: skip
                resp.raise_for_status()
                return FileUploadResponse.model_validate(resp.json())
            except httpx.TimeoutException as e:
                raise UploadTimeoutError(sandbox_id, file_path, limit) from e
            except httpx.HTTPStatusError as e:
                if e.response.status_code == 409 and self._should_retry_409(sandbox_id, e, attempt):
                    continue
                raise APIError(f"Upload failed: {http_detail(e)}") from e
            except httpx.RequestError as e:
                raise APIError(f"Upload failed: {request_detail(e)}") from e
        raise APIError("Upload failed after retries")

    def upload_file(self, sandbox_id: str, file_path: str, local_file_path: str, timeout: int | None = None) -> FileUploadResponse:
        if not os.path.exists(local_file_path):
            raise FileNotFoundError(f"Local file not found: {local_file_path}")
        return self._upload(sandbox_id, file_path, os.path.basename(local_file_path), Path(local_file_path).read_bytes(), timeout)

    def upload_bytes(self, sandbox_id: str, file_path: str, file_bytes: bytes, filename: str, timeout: int | None = None) -> FileUploadResponse:
        return self._upload(sandbox_id, file_path, filename, file_bytes, timeout)

    def download_file(self, sandbox_id: str, file_path: str, local_file_path: str, timeout: int | None = None) -> None:
        auth = self._auth_cache.get_or_refresh(sandbox_id)
        limit = timeout if timeout is not None else 300
        for attempt in range(MAX_409_RETRIES):
            try:
                resp = self._gateway_get(gateway_url(auth, "download"), headers=bearer(auth),
                                         params={"path": file_path, "sandbox_id": sandbox_id}, timeout=limit)  # fmt: skip
                resp.raise_for_status()
                if os.path.dirname(local_file_path):
                    os.makedirs(os.path.dirname(local_file_path), exist_ok=True)
                Path(local_file_path).write_bytes(resp.content)
                return
            except httpx.TimeoutException as e:
                raise DownloadTimeoutError(sandbox_id, file_path, limit) from e
            except httpx.HTTPStatusError as e:
                if e.response.status_code == 409 and self._should_retry_409(sandbox_id, e, attempt):
                    continue
                if e.response.status_code == 404:  # same message as any other failed download, but a class callers can tell apart
                    raise SandboxFileNotFoundError(f"Download failed: {http_detail(e)}") from e
                raise APIError(f"Download failed: {http_detail(e)}") from e
            except httpx.RequestError as e:
                raise APIError(f"Download failed: {request_detail(e)}") from e
        raise APIError("Download failed after retries")

    def read_file(self, sandbox_id: str, file_path: str, timeout: int | None = None) -> ReadFileResponse:
        auth = self._auth_cache.get_or_refresh(sandbox_id)
        limit = timeout if timeout is not None else 30
        for attempt in range(MAX_409_RETRIES):
            try:
                resp = self._gateway_get(gateway_url(auth, "read-file"), headers=bearer(auth), params={"path": file_path}, timeout=limit)
                resp.raise_for_status()
                return ReadFileResponse.model_validate(resp.json())
            except httpx.TimeoutException as e:
                raise APIError(f"Read file timed out after {limit}s: {file_path}") from e
            except httpx.HTTPStatusError as e:
                if e.response.status_code == 404:
                    raise SandboxFileNotFoundError(f"File not found: {file_path}") from e
                if e.response.status_code == 409 and self._should_retry_409(sandbox_id, e, attempt):
                    continue
                raise APIError(f"Read file failed: {http_detail(e)}") from e
            except httpx.RequestError as e:
                raise APIError(f"Read file failed: {request_detail(e)}") from e
        raise APIError("Read file failed after retries")

    # ---- ports / ssh
    def expose(self, sandbox_id: str, port: int, name: str | None = None, protocol: str = "HTTP") -> ExposedPort:
        body = ExposePortRequest(port=port, name=name, protocol=protocol).wire()
        return ExposedPort.model_validate(self.client.request("POST", f"/sandbox/{sandbox_id}/expose", json=body))

    def unexpose(self, sandbox_id: str, exposure_id: str) -> None:
        self.client.request("DELETE", f"/sandbox/{sandbox_id}/expose/{exposure_id}")

    def list_exposed_ports(self, sandbox_id: str) -> ListExposedPortsResponse:
        return ListExposedPortsResponse.model_validate(self.client.request("GET", f"/sandbox/{sandbox_id}/expose"))

    def list_all_exposed_ports(self) -> ListExposedPortsResponse:
        return ListExposedPortsResponse.model_validate(self.client.request("GET", "/sandbox/expose/all"))

    def create_ssh_session(self, sandbox_id: str, ttl_seconds: int | None = None) -> SSHSession:
        body = {"ttl_seconds": ttl_seconds} if ttl_seconds is not None else {}
        return SSHSession.model_validate(self.client.request("POST", f"/sandbox/{sandbox_id}/ssh-session", json=body))

    def close_ssh_session(self, sandbox_id: str, session_id: str) -> None:
        self.client.request("DELETE", f"/sandbox/{sandbox_id}/ssh-session/{session_id}")


def _list_params(client, team_id, status, labels, page, per_page, exclude_terminated) -> dict[str, Any]:
    team_id = team_id if team_id is not None else client.config.team_id
    p: dict[str, Any] = {"page": page, "per_page": per_page}
    if team_id:
        p["team_id"] = team_id
    if status:
        p["status"] = status
    if labels:
        p["labels"] = labels
    if exclude_terminated is not None:
        p["is_active"] = exclude_terminated
    return p


# --------------------------------------------------------------------------------------- async client
class AsyncSandboxClient:
    """Async mirror.  All gateway traffic shares ONE pooled ``httpx.AsyncClient`` (default 1000 connections /
    200 keep-alive) with per-request timeouts, which is what lets thousands of concurrent commands work."""

    def __init__(self, api_key: str | None = None, api_client: AsyncAPIClient | None = None, max_connections: int = 1000,
                 max_keepalive_connections: int = 200):  # fmt: skip
        self.client = api_client or AsyncAPIClient(api_key=api_key, user_agent=sandboxes_user_agent(), retry=TRANSPORT_RETRY)
        self._auth_cache = AsyncAuthCache(self.client.config.config_dir / "sandbox_auth_cache.json", self.client)
        self._limits = httpx.Limits(max_connections=max_connections, max_keepalive_connections=max_keepalive_connections)
        self._gw: httpx.AsyncClient | None = None
        self._sleep = asyncio.sleep

    def _pool(self) -> httpx.AsyncClient:
        if self._gw is None or self._gw.is_closed:
            self._gw = httpx.AsyncClient(limits=self._limits, timeout=None)
        return self._gw

    async def aclose(self) -> None:
        if self._gw is not None and not self._gw.is_closed:
            await self._gw.aclose()
        await self.client.aclose()

    async def __aenter__(self):
        return self

    async def __aexit__(self, *exc):
        await self.aclose()

    async def _gateway(self, method: str, url: str, *, idempotent: bool, headers: dict, timeout: float, **kw) -> httpx.Response:
        for attempt in range(GATEWAY_ATTEMPTS):
            try:
                resp = await self._pool().request(method, url, headers=headers, timeout=timeout, **kw)
                if idempotent and resp.status_code in RETRYABLE_5XX:
                    resp.raise_for_status()
                return resp
            except Exception as e:
                if attempt + 1 < GATEWAY_ATTEMPTS and retryable(e, idempotent):
                    await self._sleep(backoff(attempt))
                    continue
                raise
        raise AssertionError("unreachable")

    async def _gateway_post(self, url: str, headers: dict, timeout: float, **kw) -> httpx.Response:
        return await self._gateway("POST", url, idempotent=False, headers=headers, timeout=timeout, **kw)

    async def _gateway_get(self, url: str, headers: dict, params: dict, timeout: float) -> httpx.Response:
        return await self._gateway("GET", url, idempotent=True, headers=headers, timeout=timeout, params=params)

    async def _get_sandbox_error_context(self, sandbox_id: str) -> dict:
        try:
            return parse_error_context(await self.client.request("GET", f"/sandbox/{sandbox_id}/error-context"))
        except Exception:
            return dict(EMPTY_CTX)

    async def _should_retry_409(self, sandbox_id: str, error: httpx.HTTPStatusError, attempt: int, command: str | None = None) -> bool:
        ctx = await self._get_sandbox_error_context(sandbox_id)
        if ctx["status"] != "RUNNING":
            classify_not_running(sandbox_id, ctx, command=command, cause=error)
        if attempt < MAX_409_RETRIES - 1:
            await self._sleep(RETRY_409_BASE_DELAY * (2**attempt))
            return True
        raise APIError(f"Sandbox {sandbox_id} returned 409 after {MAX_409_RETRIES} retries. "
                       "This may be a transient DNS or gateway issue. Please retry.") from error  # fmt: skip

    async def _timeout_or_dead(self, sandbox_id: str, command: str, timeout: int, cause: BaseException) -> NoReturn:
        ctx = await self._get_sandbox_error_context(sandbox_id)
        if ctx["status"] in SandboxStatus.terminal():
            classify_not_running(sandbox_id, ctx, command=command, cause=cause)
        raise CommandTimeoutError(sandbox_id, command, timeout) from cause

    async def _is_sandbox_reachable(self, sandbox_id: str, timeout: int = 10) -> bool:
        try:
            await self.execute_command(sandbox_id, "echo 'sandbox ready'", timeout=timeout)
            return True
        except Exception:
            return False

    async def clear_auth_cache(self) -> None:
        await self._auth_cache.clear()

    async def create(self, request: CreateSandboxRequest) -> Sandbox:
        if request.team_id is None:
            request.team_id = self.client.config.team_id
        return Sandbox.model_validate(await self.client.request("POST", "/sandbox", json=request.wire()))

    async def list(self, team_id: str | None = None, status: str | None = None, labels: list[str] | None = None, page: int = 1,
                   per_page: int = 50, exclude_terminated: bool | None = None) -> SandboxListResponse:  # fmt: skip
        params = _list_params(self.client, team_id, status, labels, page, per_page, exclude_terminated)
        return SandboxListResponse.model_validate(await self.client.request("GET", "/sandbox", params=params))

    async def get(self, sandbox_id: str) -> Sandbox:
        return Sandbox.model_validate(await self.client.request("GET", f"/sandbox/{sandbox_id}"))

    async def delete(self, sandbox_id: str) -> dict[str, Any]:
        return await self.client.request("DELETE", f"/sandbox/{sandbox_id}")

    async def bulk_delete(self, sandbox_ids: list[str] | None = None, labels: list[str] | None = None) -> BulkDeleteSandboxResponse:
        body = BulkDeleteSandboxRequest(sandbox_ids=sandbox_ids, labels=labels).wire()
        return BulkDeleteSandboxResponse.model_validate(await self.client.request("DELETE", "/sandbox", json=body))

    async def get_logs(self, sandbox_id: str) -> str:
        return SandboxLogsResponse.model_validate(await self.client.request("GET", f"/sandbox/{sandbox_id}/logs")).logs

    async def execute_command(self, sandbox_id: str, command: str, working_dir: str | None = None,
                              env: dict[str, str] | None = None, timeout: int | None = None) -> CommandResponse:  # fmt: skip
        auth = await self._auth_cache.get_or_refresh(sandbox_id)
        vm = await self._auth_cache.is_vm(sandbox_id)
        run = self._execute_command_connect_rpc if vm else self._execute_command_rest
        return await run(sandbox_id=sandbox_id, command=command, auth=auth, working_dir=working_dir, env=env, timeout=timeout)

    async def _execute_command_connect_rpc(self, sandbox_id: str, command: str, auth: dict, working_dir: str | None = None,
                                           env: dict[str, str] | None = None, timeout: int | None = None) -> CommandResponse:  # fmt: skip
        from connectrpc.code import Code
        from connectrpc.errors import ConnectError

        from .rpc_command_session import START_METHOD, OutputCollector, build_start_request

        limit = timeout if timeout is not None else DEFAULT_COMMAND_TIMEOUT
        rpc = _rpc_class("ConnectClient")(gateway_url(auth))
        out = OutputCollector()
        try:
            async for ev in rpc.execute_server_stream(request=build_start_request(command, working_dir, env), method=START_METHOD,
                                                      headers=bearer(auth), timeout_ms=limit * 1000):  # fmt: skip
                out.feed(ev)
            stdout, stderr, code = out.result()
            if code is None:
                raise APIError("Command stream ended without exit code")
            return CommandResponse(stdout=stdout, stderr=stderr, exit_code=code)
        except ConnectError as e:
            if e.code == Code.DEADLINE_EXCEEDED:
                await self._timeout_or_dead(sandbox_id, command, limit, e)
            if e.code == Code.NOT_FOUND:
                classify_not_running(sandbox_id, mark_gone(await self._get_sandbox_error_context(sandbox_id)), command=command, cause=e)
            raise APIError(f"Connect RPC failed ({e.code.value}): {e.message}") from e
        except (APIError, SandboxNotRunningError, CommandTimeoutError):
            raise
        except Exception as e:
            raise APIError(f"Request failed: {type(e).__name__}: {e}") from e
        finally:
            await rpc.close()

    async def _execute_command_rest(self, sandbox_id: str, command: str, auth: dict, working_dir: str | None = None,
                                    env: dict[str, str] | None = None, timeout: int | None = None) -> CommandResponse:  # fmt: skip
        limit = timeout if timeout is not None else DEFAULT_COMMAND_TIMEOUT
        url, payload = gateway_url(auth, "exec"), exec_payload(sandbox_id, command, working_dir, env, limit)
        for attempt in range(MAX_409_RETRIES):
            try:
                resp = await self._gateway_post(url, headers=bearer(auth), timeout=limit + 5, json=payload)
                resp.raise_for_status()
                return CommandResponse.model_validate(resp.json())
            except httpx.TimeoutException as e:
                await self._timeout_or_dead(sandbox_id, command, limit, e)
            except httpx.HTTPStatusError as e:
                code = e.response.status_code
                if gateway_says_sandbox_gone(e.response):
                    classify_not_running(sandbox_id, mark_gone(await self._get_sandbox_error_context(sandbox_id)), command=command, cause=e)
                if code == 409 and await self._should_retry_409(sandbox_id, e, attempt, command=command):
                    continue
                if code == 408:
                    await self._timeout_or_dead(sandbox_id, command, limit, e)
                raise APIError(http_detail(e)) from e
            except httpx.RequestError as e:
                raise APIError(f"Request failed: {request_detail(e)}") from e
        raise APIError("Command execution failed after retries")

    async def start_background_job(self, sandbox_id: str, command: str, working_dir: str | None = None,
                                   env: dict[str, str] | None = None) -> BackgroundJob:  # fmt: skip
        job, shell = background_command(command, working_dir, env)
        job.sandbox_id = sandbox_id
        await self.execute_command(sandbox_id, shell, timeout=10)
        return job

    async def _read_or_empty(self, sandbox_id: str, path: str) -> str:
        try:
            return (await self.read_file(sandbox_id, path)).content
        except SandboxFileNotFoundError:
            return ""

    async def get_background_job(self, sandbox_id: str, job: BackgroundJob) -> BackgroundJobStatus:
        code = parse_exit_marker(await self._read_or_empty(sandbox_id, job.exit_file))
        if code is None:
            return BackgroundJobStatus(job_id=job.job_id, completed=False)
        return BackgroundJobStatus(job_id=job.job_id, completed=True, exit_code=code,
                                   stdout=await self._read_or_empty(sandbox_id, job.stdout_log_file),
                                   stderr=await self._read_or_empty(sandbox_id, job.stderr_log_file))  # fmt: skip

    async def run_background_job(self, sandbox_id: str, command: str, timeout: int = 900, working_dir: str | None = None,
                                 env: dict[str, str] | None = None, poll_interval: int = 3) -> BackgroundJobStatus:  # fmt: skip
        job = await self.start_background_job(sandbox_id, command, working_dir=working_dir, env=env)
        deadline = time.monotonic() + timeout
        while time.monotonic() < deadline:
            st = await self.get_background_job(sandbox_id, job)
            if st.completed:
                return st
            await self._sleep(poll_interval)
        raise CommandTimeoutError(sandbox_id, command, timeout)

    async def wait_for_creation(self, sandbox_id: str, max_attempts: int = 60, stability_checks: int = 1) -> None:
        streak = 0
        for attempt in range(max_attempts):
            sb = await self.get(sandbox_id)
            if sb.status == "RUNNING":
                if await self._is_sandbox_reachable(sandbox_id):
                    streak += 1
                    if streak >= stability_checks:
                        return
                    await self._sleep(0.5)
                    continue
                streak = 0
            elif sb.status in SandboxStatus.terminal():
                classify_not_running(sb.id, {"status": sb.status, "error_type": sb.error_type, "error_message": sb.error_message})
            await self._sleep(1 if attempt < 5 else 2)
        raise SandboxNotRunningError(sandbox_id, "Timeout during sandbox creation")

    async def bulk_wait_for_creation(self, sandbox_ids: list[str], max_attempts: int = 60) -> dict[str, str]:
        wanted, final = set(sandbox_ids), {}
        for attempt in range(max_attempts):
            running, failed, page = 0, [], 1
            while True:
                try:
                    resp = await self.list(per_page=100, page=page)
                except Exception as e:
                    if is_rate_limited(e):
                        await self._sleep(min(2**attempt, 60))
                        continue
                    raise
                r, f, st = tally_statuses(resp.sandboxes, wanted)
                running, failed = running + r, failed + f
                final.update(st)
                if len(final) == len(sandbox_ids) or not resp.has_next:
                    break
                page += 1
            if failed:
                raise RuntimeError(f"Sandboxes failed: {failed}")
            if running == len(sandbox_ids):
                checks = await asyncio.gather(*(self._is_sandbox_reachable(s) for s in sandbox_ids))
                unreachable = [s for s, ok in zip(sandbox_ids, checks) if not ok]
                for s in unreachable:
                    final.pop(s, None)
                if not unreachable:
                    return final
            await self._sleep(1 if attempt < 5 else 2)
        for s in wanted - set(final):
            final[s] = "TIMEOUT"
        raise RuntimeError(f"Timeout waiting for sandboxes to be ready. Status: {final}")

    async def _upload(self, sandbox_id: str, file_path: str, filename: str, content: bytes, timeout: int | None) -> FileUploadResponse:
        auth = await self._auth_cache.get_or_refresh(sandbox_id)
        limit = timeout if timeout is not None else 300
        for attempt in range(MAX_409_RETRIES):
            try:
                resp = await self._gateway_post(gateway_url(auth, "upload"), headers=bearer(auth), timeout=limit,
                                                files={"file": (filename, content)},
                                                params={"path": file_path, "sandbox_id": sandbox_id})  # fmt: skip
                resp.raise_for_status()
                return FileUploadResponse.model_validate(resp.json())
            except httpx.TimeoutException as e:
                raise UploadTimeoutError(sandbox_id, file_path, limit) from e
            except httpx.HTTPStatusError as e:
                if e.response.status_code == 409 and await self._should_retry_409(sandbox_id, e, attempt):
                    continue
                raise APIError(f"Upload failed: {http_detail(e)}") from e
            except httpx.RequestError as e:
                raise APIError(f"Upload failed: {request_detail(e)}") from e
        raise APIError("Upload failed after retries")

    async def upload_file(self, sandbox_id: str, file_path: str, local_file_path: str, timeout: int | None = None) -> FileUploadResponse:
        if not os.path.exists(local_file_path):
            raise FileNotFoundError(f"Local file not found: {local_file_path}")
        import aiofiles

        async with aiofiles.open(local_file_path, "rb") as f:
            content = await f.read()
        return await self._upload(sandbox_id, file_path, os.path.basename(local_file_path), content, timeout)

    async def upload_bytes(self, sandbox_id: str, file_path: str, file_bytes: bytes, filename: str, timeout: int | None = None) -> FileUploadResponse:
        return await self._upload(sandbox_id, file_path, filename, file_bytes, timeout)

    async def download_file(self, sandbox_id: str, file_path: str, local_file_path: str, timeout: int | None = None) -> None:
        auth = await self._auth_cache.get_or_refresh(sandbox_id)
        limit = timeout if timeout is not None else 300
        for attempt in range(MAX_409_RETRIES):
            try:
                resp = await self._gateway_get(gateway_url(auth, "download"), headers=bearer(auth),
                                               params={"path": file_path, "sandbox_id": sandbox_id}, timeout=limit)  # fmt: skip
                resp.raise_for_status()
                if os.path.dirname(local_file_path):
                    os.makedirs(os.path.dirname(local_file_path), exist_ok=True)
                import aiofiles

                async with aiofiles.open(local_file_path, "wb") as f:
                    await f.write(resp.content)
                return
            except httpx.TimeoutException as e:
                raise DownloadTimeoutError(sandbox_id, file_path, limit) from e
            except httpx.HTTPStatusError as e:
                if e.response.status_code == 409 and await self._should_retry_409(sandbox_id, e, attempt):
                    continue
                if e.response.status_code == 404:  # same message as any other failed download, but a class callers can tell apart
                    raise SandboxFileNotFoundError(f"Download failed: {http_detail(e)}") from e
                raise APIError(f"Download failed: {http_detail(e)}") from e
            except httpx.RequestError as e:
                raise APIError(f"Download failed: {request_detail(e)}") from e
        raise APIError("Download failed after retries")

    async def read_file(self, sandbox_id: str, file_path: str, timeout: int | None = None) -> ReadFileResponse:
        auth = await self._auth_cache.get_or_refresh(sandbox_id)
        limit = timeout if timeout is not None else 30
        for attempt in range(MAX_409_RETRIES):
            try:
                resp = await self._gateway_get(gateway_url(auth, "read-file"), headers=bearer(auth), params={"path": file_path}, timeout=limit)
                resp.raise_for_status()
                return ReadFileResponse.model_validate(resp.json())
            except httpx.TimeoutException as e:
                raise APIError(f"Read file timed out after {limit}s: {file_path}") from e
            except httpx.HTTPStatusError as e:
                if e.response.status_code == 404:
                    raise SandboxFileNotFoundError(f"File not found: {file_path}") from e
                if e.response.status_code == 409 and await self._should_retry_409(sandbox_id, e, attempt):
                    continue
                raise APIError(f"Read file failed: {http_detail(e)}") from e
            except httpx.RequestError as e:
                raise APIError(f"Read file failed: {request_detail(e)}") from e
        raise APIError("Read file failed after retries")

    async def expose(self, sandbox_id: str, port: int, name: str | None = None, protocol: str = "HTTP") -> ExposedPort:
        body = ExposePortRequest(port=port, name=name, protocol=protocol).wire()
        return ExposedPort.model_validate(await self.client.request("POST", f"/sandbox/{sandbox_id}/expose", json=body))

    async def unexpose(self, sandbox_id: str, exposure_id: str) -> None:
        await self.client.request("DELETE", f"/sandbox/{sandbox_id}/expose/{exposure_id}")

    async def list_exposed_ports(self, sandbox_id: str) -> ListExposedPortsResponse:
        return ListExposedPortsResponse.model_validate(await self.client.request("GET", f"/sandbox/{sandbox_id}/expose"))

    async def list_all_exposed_ports(self) -> ListExposedPortsResponse:
        return ListExposedPortsResponse.model_validate(await self.client.request("GET", "/sandbox/expose/all"))

    async def create_ssh_session(self, sandbox_id: str, ttl_seconds: int | None = None) -> SSHSession:
        body = {"ttl_seconds": ttl_seconds} if ttl_seconds is not None else {}
        return SSHSession.model_validate(await self.client.request("POST", f"/sandbox/{sandbox_id}/ssh-session", json=body))

    async def close_ssh_session(self, sandbox_id: str, session_id: str) -> None:
        await self.client.request("DELETE", f"/sandbox/{sandbox_id}/ssh-session/{session_id}")


# --------------------------------------------------------------------------------------- templates
class TemplateClient:
    """Registry credentials + docker image accessibility checks."""

    def __init__(self, api_client: APIClient | None = None):
        self.client = api_client or APIClient(user_agent=sandboxes_user_agent(), retry=TRANSPORT_RETRY)

    def list_registry_credentials(self) -> list[RegistryCredentialSummary]:
        resp = self.client.request("GET", "/template/registry-credentials")
        return [RegistryCredentialSummary.model_validate(x) for x in resp.get("credentials", resp.get("data", []))]

    def check_docker_image(self, image: str, registry_credentials_id: str | None = None) -> DockerImageCheckResponse:
        body: dict[str, Any] = {"image": image}
        if registry_credentials_id:
            body["registry_credentials_id"] = registry_credentials_id
        return DockerImageCheckResponse.model_validate(self.client.request("POST", "/template/check-docker-image", json=body))


class AsyncTemplateClient:
    def __init__(self, api_client: AsyncAPIClient | None = None):
        self.client = api_client or AsyncAPIClient(user_agent=sandboxes_user_agent(), retry=TRANSPORT_RETRY)

    async def aclose(self) -> None:
        await self.client.aclose()

    async def __aenter__(self):
        return self

    async def __aexit__(self, *exc):
        await self.aclose()

    async def list_registry_credentials(self) -> list[RegistryCredentialSummary]:
        resp = await self.client.request("GET", "/template/registry-credentials")
        return [RegistryCredentialSummary.model_validate(x) for x in resp.get("credentials", resp.get("data", []))]

    async def check_docker_image(self, image: str, registry_credentials_id: str | None = None) -> DockerImageCheckResponse:
        body: dict[str, Any] = {"image": image}
        if registry_credentials_id:
            body["registry_credentials_id"] = registry_credentials_id
        return DockerImageCheckResponse.model_validate(await self.client.request("POST", "/template/check-docker-image", json=body))
