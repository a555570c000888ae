"""Sandbox SDK behaviour with substituted transports (no network), following the reference's approach
(prime-sandboxes/tests/test_client_retry.py, test_command_transport_selection.py, test_gateway_error_mapping.py)."""

import json
import threading
from datetime import datetime, timedelta, timezone

import httpx
import pytest

from prime_b200.platform.core import APIClient, APIError
from prime_b200.platform.sandboxes import (
    CommandTimeoutError,
    CreateSandboxRequest,
    SandboxClient,
    SandboxFileNotFoundError,
    SandboxNotRunningError,
    SandboxOOMError,
)
from prime_b200.platform.sandboxes import sandbox as sb

FUTURE = (datetime.now(timezone.utc) + timedelta(hours=1)).isoformat()
PAST = (datetime.now(timezone.utc) - timedelta(hours=1)).isoformat()
AUTH = {"gateway_url": "https://gw.example", "user_ns": "ns", "job_id": "job", "token": "tok", "expires_at": FUTURE, "is_vm": False}


class FakeAPI:
    """Stands in for APIClient: records control-plane calls."""

    def __init__(self, home, routes=None):
        from prime_b200.platform.core import Config

        self.config = Config(writable=False)
        self.calls = []
        self.routes = routes or {}

    def request(self, method, endpoint, params=None, json=None, timeout=None):
        self.calls.append((method, endpoint))
        r = self.routes.get((method, endpoint))
        if callable(r):
            return r()
        if r is None:
            raise APIError(f"no route {method} {endpoint}")
        return r


def make_client(isolated_home, routes=None, gateway=None):
    api = FakeAPI(isolated_home, routes or {("POST", "/sandbox/s1/auth"): dict(AUTH)})
    c = SandboxClient(api)
    c._sleep = lambda s: None
    if gateway is not None:
        c._gateway = gateway
    return c, api


def resp(code, body=None, url="https://gw.example/ns/job/exec", method="POST"):
    r = httpx.Response(code, json=body if body is not None else {}, request=httpx.Request(method, url))
    return r


def test_create_request_gpu_validation():
    with pytest.raises(ValueError, match="gpu_type is required"):
        CreateSandboxRequest(name="a", docker_image="x", gpu_count=1, vm=True)
    with pytest.raises(ValueError, match="only supported when vm"):
        CreateSandboxRequest(name="a", docker_image="x", gpu_count=1, gpu_type="H100")
    with pytest.raises(ValueError, match="requires gpu_count"):
        CreateSandboxRequest(name="a", docker_image="x", gpu_type="H100")
    ok = CreateSandboxRequest(name="a", docker_image="x", gpu_count=1, gpu_type="H100", vm=True)
    assert ok.wire()["gpu_type"] == "H100" and "team_id" not in ok.wire()


def test_auth_cache_prunes_expired_and_persists(isolated_home):
    path = isolated_home / ".prime" / "sandbox_auth_cache.json"
    path.parent.mkdir(parents=True)
    path.write_text(json.dumps({"old": {**AUTH, "expires_at": PAST}, "new": dict(AUTH)}))
    api = FakeAPI(isolated_home)
    cache = sb.AuthCache(path, api)
    assert set(json.loads(path.read_text())) == {"new"}
    assert cache.get_or_refresh("new")["token"] == "tok" and api.calls == []


def test_auth_single_flight(isolated_home):
    """N concurrent misses → exactly one POST /auth (reference: sandbox.py:263-299)."""
    gate = threading.Event()
    n_posts = []

    def slow_auth():
        n_posts.append(1)
        gate.wait(2)
        return dict(AUTH)

    api = FakeAPI(isolated_home, {("POST", "/sandbox/s1/auth"): slow_auth})
    cache = sb.AuthCache(isolated_home / ".prime" / "c.json", api)
    out = []
    ts = [threading.Thread(target=lambda: out.append(cache.get_or_refresh("s1")["token"])) for _ in range(8)]
    for t in ts:
        t.start()
    import time

    time.sleep(0.2)
    gate.set()
    for t in ts:
        t.join(5)
    assert out == ["tok"] * 8 and len(n_posts) == 1


def test_transport_selection_vm_vs_container(isolated_home):
    c, api = make_client(isolated_home)
    c._execute_command_rest = lambda **kw: "rest"
    c._execute_command_connect_rpc = lambda **kw: "rpc"
    assert c.execute_command("s1", "ls") == "rest"
    c._auth_cache.set("s1", {**AUTH, "is_vm": True})
    assert c.execute_command("s1", "ls") == "rpc"


def test_rest_exec_success_and_payload(isolated_home):
    seen = {}

    def gw(method, url, *, idempotent, headers, timeout, **kw):
        seen.update(method=method, url=url, idem=idempotent, auth=headers["Authorization"], timeout=timeout, json=kw["json"])
        return resp(200, {"stdout": "hi\n", "stderr": "", "exit_code": 0})

    c, _ = make_client(isolated_home, gateway=gw)
    r = c.execute_command("s1", "echo hi", working_dir="/w", env={"A": "1"}, timeout=7)
    assert r.stdout == "hi\n" and r.exit_code == 0
    assert seen["url"] == "https://gw.example/ns/job/exec" and seen["idem"] is False and seen["auth"] == "Bearer tok"
    assert seen["timeout"] == 12 and seen["json"]["timeout"] == 7 and seen["json"]["working_dir"] == "/w"


def test_409_retries_only_while_running(isolated_home):
    calls = []

    def gw(method, url, **kw):
        calls.append(1)
        if len(calls) < 3:
            return resp(409, {"detail": "busy"})
        return resp(200, {"stdout": "", "stderr": "", "exit_code": 0})

    routes = {("POST", "/sandbox/s1/auth"): dict(AUTH), ("GET", "/sandbox/s1/error-context"): {"status": "RUNNING"}}
    c, _ = make_client(isolated_home, routes, gw)
    assert c.execute_command("s1", "x").exit_code == 0 and len(calls) == 3

    calls.clear()
    routes[("GET", "/sandbox/s1/error-context")] = {"status": "TERMINATED", "errorType": "OOM_KILLED", "errorMessage": "boom"}
    c, _ = make_client(isolated_home, routes, lambda *a, **k: (calls.append(1), resp(409))[1])
    with pytest.raises(SandboxOOMError, match="out-of-memory"):
        c.execute_command("s1", "x")
    assert len(calls) == 1

    calls.clear()
    routes[("GET", "/sandbox/s1/error-context")] = {"status": "RUNNING"}
    c, _ = make_client(isolated_home, routes, lambda *a, **k: (calls.append(1), resp(409))[1])
    with pytest.raises(APIError, match="409 after 4 retries"):
        c.execute_command("s1", "x")
    assert len(calls) == 4


def test_gateway_sandbox_not_found_maps_to_not_running(isolated_home):
    routes = {("POST", "/sandbox/s1/auth"): dict(AUTH)}
    c, _ = make_client(isolated_home, routes, lambda *a, **k: resp(502, {"error": "sandbox_not_found"}))
    with pytest.raises(SandboxNotRunningError) as e:
        c.execute_command("s1", "python train.py")
    assert e.value.error_type == "SANDBOX_NOT_FOUND" and e.value.status == "TERMINATED"
    assert "no longer present" in str(e.value)


def test_timeout_vs_dead_sandbox(isolated_home):
    def gw(method, url, **kw):
        raise httpx.ReadTimeout("t", request=httpx.Request("POST", url))

    routes = {("POST", "/sandbox/s1/auth"): dict(AUTH), ("GET", "/sandbox/s1/error-context"): {"status": "RUNNING"}}
    c, _ = make_client(isolated_home, routes, gw)
    with pytest.raises(CommandTimeoutError, match="timed out after 5s"):
        c.execute_command("s1", "sleep 99", timeout=5)
    routes[("GET", "/sandbox/s1/error-context")] = {"status": "TIMEOUT", "error_type": "TIMEOUT"}
    with pytest.raises(SandboxNotRunningError, match="maximum runtime"):
        c.execute_command("s1", "sleep 99", timeout=5)


def test_retry_split_by_idempotency(isolated_home, monkeypatch):
    """GET retries 5xx, POST retries only connection-level failures (reference: sandbox.py:94-109)."""
    c, _ = make_client(isolated_home)
    calls = []

    built = []

    class FakeHttpxClient:
        is_closed = False

        def __init__(self, timeout=None, limits=None):
            built.append(self)

        def request(self, method, url, headers=None, **kw):
            assert kw.pop("timeout") == 1  # the pooled client carries no default: every request brings its own limit
            calls.append(method)
            if len(calls) < 3:
                return resp(503, {"detail": "x"}, url=url, method=method)
            return resp(200, {"content": "ok", "size": 2}, url=url, method=method)

    monkeypatch.setattr(sb.httpx, "Client", FakeHttpxClient)
    r = c._gateway_get("https://gw.example/ns/job/read-file", headers={}, params={}, timeout=1)
    assert r.status_code == 200 and len(calls) == 3
    calls.clear()
    r = c._gateway_post("https://gw.example/ns/job/exec", headers={}, timeout=1, json={})
    assert r.status_code == 503 and len(calls) == 1  # not retried: the server saw it
    assert len(built) == 1  # one pooled gateway client for all of it (the reference builds one per request)


def test_retryable_predicate():
    req = httpx.Request("GET", "https://x")
    assert sb.retryable(httpx.ConnectError("x", request=req), idempotent=False)
    assert not sb.retryable(httpx.ReadTimeout("x", request=req), idempotent=True)
    e503 = httpx.HTTPStatusError("x", request=req, response=httpx.Response(503, request=req))
    assert sb.retryable(e503, True) and not sb.retryable(e503, False)
    gone = httpx.HTTPStatusError("x", request=req, response=httpx.Response(502, json={"error": "sandbox_not_found"}, request=req))
    assert not sb.retryable(gone, True)


def test_background_job_wrapping_and_polling(isolated_home):
    import subprocess
    import time

    wd = isolated_home / "work dir"
    wd.mkdir()
    job, shell = sb.background_command('echo "$K in $(pwd)"; exit 3', str(wd), {"K": "v w"})
    assert shell.startswith("nohup sh -c ") and shell.endswith("&")
    subprocess.run(["bash", "-c", shell], check=True)  # the wrapper really works in a shell
    for _ in range(50):
        if open(job.exit_file).read().strip() if __import__("os").path.exists(job.exit_file) else "":
            break
        time.sleep(0.05)
    assert open(job.exit_file).read().strip() == "3"  # `exit` inside the subshell cannot skip the marker
    assert open(job.stdout_log_file).read().strip() == f"v w in {wd}"
    with pytest.raises(ValueError):
        sb.background_command("x", None, {"bad-key": "1"})
    assert sb.parse_exit_marker("") is None and sb.parse_exit_marker("3\n") == 3 and sb.parse_exit_marker("x") is None

    c, _ = make_client(isolated_home)
    files = {}
    c.execute_command = lambda sid, cmd, **kw: None

    def read_file(sid, path, timeout=None):
        if path not in files:
            raise SandboxFileNotFoundError("nf")
        from prime_b200.platform.sandboxes import ReadFileResponse

        return ReadFileResponse(content=files[path], size=len(files[path]))

    c.read_file = read_file
    j = c.start_background_job("s1", "make")
    assert not c.get_background_job("s1", j).completed
    files[j.exit_file], files[j.stdout_log_file] = "0\n", "built"
    st = c.get_background_job("s1", j)
    assert st.completed and st.exit_code == 0 and st.stdout == "built" and st.stderr == ""


def test_read_file_404(isolated_home):
    c, _ = make_client(isolated_home, gateway=lambda *a, **k: resp(404, {}, method="GET"))
    with pytest.raises(SandboxFileNotFoundError):
        c.read_file("s1", "/nope")


def test_download_of_a_missing_file_is_a_file_not_found_error_with_the_download_message(isolated_home, tmp_path):
    """404 on download: the reference raises a bare APIError("Download failed: …"); here the same message on SandboxFileNotFoundError (an
    APIError subclass — `except APIError` keeps working), so a caller can tell a missing file from a broken gateway. 500 stays APIError."""
    c, _ = make_client(isolated_home, gateway=lambda *a, **k: resp(404, {"detail": "file not found: /nope"}, method="GET"))
    with pytest.raises(SandboxFileNotFoundError, match="Download failed: HTTP 404") as e:
        c.download_file("s1", "/nope", str(tmp_path / "x"))
    assert isinstance(e.value, APIError) and not (tmp_path / "x").exists()
    c, _ = make_client(isolated_home, gateway=lambda *a, **k: resp(500, {}, method="GET"))
    with pytest.raises(APIError) as e:
        c.download_file("s1", "/nope", str(tmp_path / "x"))
    assert not isinstance(e.value, SandboxFileNotFoundError)


@pytest.mark.anyio
async def test_async_download_of_a_missing_file(isolated_home, tmp_path):
    from prime_b200.platform.sandboxes import AsyncSandboxClient

    class FakeAsyncAPI:
        def __init__(self):
            from prime_b200.platform.core import Config

            self.config = Config(writable=False)

        async def request(self, method, endpoint, params=None, json=None, timeout=None):
            return dict(AUTH)

        async def aclose(self):
            pass

    c = AsyncSandboxClient(api_client=FakeAsyncAPI())

    async def gw(method, url, *, idempotent, headers, timeout, **kw):
        return resp(404, {"detail": "file not found"}, method="GET")

    c._gateway = gw
    with pytest.raises(SandboxFileNotFoundError, match="Download failed"):
        await c.download_file("s1", "/nope", str(tmp_path / "x"))
    await c.aclose()


def test_wait_for_creation_failure_classification(isolated_home):
    sbx = {"id": "s1", "name": "n", "dockerImage": "i", "cpuCores": 1, "memoryGB": 1, "diskSizeGB": 1, "diskMountPath": "/", "gpuCount": 0,
           "status": "ERROR", "timeoutMinutes": 1, "createdAt": FUTURE, "updatedAt": FUTURE, "errorType": "IMAGE_PULL_FAILED",
           "errorMessage": "manifest unknown"}
    c, _ = make_client(isolated_home, {("GET", "/sandbox/s1"): sbx})
    from prime_b200.platform.sandboxes import SandboxImagePullError

    with pytest.raises(SandboxImagePullError, match="manifest unknown"):
        c.wait_for_creation("s1")


def test_tally_statuses():
    from prime_b200.platform.sandboxes import Sandbox

    def mk(i, st):
        return Sandbox.model_validate({"id": i, "name": i, "dockerImage": "i", "cpuCores": 1, "memoryGB": 1, "diskSizeGB": 1,
                                       "diskMountPath": "/", "gpuCount": 0, "status": st, "timeoutMinutes": 1,
                                       "createdAt": FUTURE, "updatedAt": FUTURE})

    running, failed, st = sb.tally_statuses([mk("a", "RUNNING"), mk("b", "ERROR"), mk("c", "PENDING"), mk("z", "RUNNING")], {"a", "b", "c"})
    assert running == 1 and failed == [("b", "ERROR")] and st == {"a": "RUNNING", "b": "ERROR", "c": "PENDING"}


@pytest.mark.anyio
async def test_async_exec_and_single_flight(isolated_home):
    import asyncio

    from prime_b200.platform.sandboxes import AsyncSandboxClient

    posts = []

    class FakeAsyncAPI:
        def __init__(self):
            from prime_b200.platform.core import Config

            self.config = Config(writable=False)

        async def request(self, method, endpoint, params=None, json=None, timeout=None):
            if endpoint.endswith("/auth"):
                posts.append(1)
                await asyncio.sleep(0.05)
                return dict(AUTH)
            raise APIError("no route")

        async def aclose(self):
            pass

    c = AsyncSandboxClient(api_client=FakeAsyncAPI())

    async def gw(method, url, *, idempotent, headers, timeout, **kw):
        return resp(200, {"stdout": kw["json"]["command"], "stderr": "", "exit_code": 0})

    c._gateway = gw
    outs = await asyncio.gather(*(c.execute_command("s1", f"c{i}") for i in range(20)))
    assert [o.stdout for o in outs] == [f"c{i}" for i in range(20)] and len(posts) == 1
    await c.aclose()
