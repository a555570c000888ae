#!/usr/bin/env python
"""Head-to-head of the platform layer against the UNMODIFIED reference packages, on the CPU, where the reference CAN run.

    python tools/platform_bench.py > profiles/platform_bench.json

The reference (`/root/reference/packages/*/src`, imported from its source tree, not installed or modified) and
`prime_b200.platform` are driven through their public APIs against the same local fake control plane + gateway
(`ThreadingHTTPServer`, HTTP/1.1 keep-alive) in separate child processes with a private HOME:

* ``cli_cold_start``   — wall time of ``<cli> --help`` and ``<cli> sandbox --help`` in a fresh interpreter (median of 7);
* ``sandbox_exec``     — ``SandboxClient.execute_command`` (the SDK's hot path): sequential calls/s on one client and
                         with 16 threads sharing one client (auth cached after the first call, as in real use);
* ``async_fanout``     — ``AsyncSandboxClient.execute_command``: 2000 calls under ``Semaphore(64)`` (the reference's high-volume demo);
* ``eval_push``        — ``EvalsClient.push_samples`` of 20 000 samples (batching to 2 MiB + 4-way upload).

Both arms talk to the same server process; the server's own cost is a floor common to both, so the ratios UNDER-state
differences in client overhead. Every number is a median over repetitions, wall clock (``time.perf_counter``).
"""

from __future__ import annotations

import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/packages")
REF_PATHS = [str(REF / p / "src") for p in ("prime", "prime-evals", "prime-sandboxes", "prime-tunnel", "prime-mcp-server")]
FUTURE = "2099-01-01T00:00:00Z"
SANDBOX = {"id": "s1", "name": "bench", "dockerImage": "python:3.11-slim", "startCommand": None, "cpuCores": 2, "memoryGB": 4, "diskSizeGB": 10,
           "diskMountPath": "/workspace", "gpuCount": 0, "gpuType": None, "vm": False, "status": "RUNNING", "timeoutMinutes": 60,
           "createdAt": "2025-01-01T00:00:00Z", "updatedAt": "2025-01-01T00:00:00Z", "userId": "u1", "teamId": None}  # fmt: skip


class Handler(BaseHTTPRequestHandler):
    protocol_version = "HTTP/1.1"
    counts: dict[str, int] = {}

    def log_message(self, *a):  # noqa: D102
        pass

    def setup(self):
        import socket

        super().setup()
        self.request.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)

    def _send(self, obj, code=200):
        # one write per response (status line + headers + body): two writes would meet the client's delayed ACK (40 ms per call)
        body = json.dumps(obj).encode()
        head = f"HTTP/1.1 {code} {'OK' if code == 200 else 'ERR'}\r\nContent-Type: application/json\r\nContent-Length: {len(body)}\r\n\r\n".encode()
        self.wfile.write(head + body)
        self.wfile.flush()

    def _body(self):
        n = int(self.headers.get("Content-Length") or 0)
        return self.rfile.read(n) if n else b""

    def do_GET(self):  # noqa: N802
        p = self.path.split("?")[0]
        Handler.counts[p] = Handler.counts.get(p, 0) + 1
        if p.endswith("/sandbox/s1"):
            return self._send(SANDBOX)
        if p.endswith("/error-context"):
            return self._send({"status": "RUNNING"})
        if p.endswith("/__counts"):
            return self._send(Handler.counts)
        self._send({"detail": "not found"}, 404)

    def do_POST(self):  # noqa: N802
        p = self.path.split("?")[0]
        raw = self._body()
        Handler.counts[p] = Handler.counts.get(p, 0) + 1
        host = self.headers.get("Host")
        if p.endswith("/sandbox/s1/auth"):
            return self._send({"gateway_url": f"http://{host}/gw", "user_ns": "ns", "job_id": "job", "token": "tok", "expires_at": FUTURE, "is_vm": False})
        if p == "/gw/ns/job/exec":
            return self._send({"stdout": "ok\n", "stderr": "", "exit_code": 0})
        if p.endswith("/samples"):
            Handler.counts["sample_bytes"] = Handler.counts.get("sample_bytes", 0) + len(raw)
            return self._send({"status": "ok", "count": len(json.loads(raw).get("samples", []))})
        if "/evaluations" in p:
            return self._send({"evaluation_id": "ev1", "id": "ev1", "status": "ok"})
        self._send({"detail": "not found"}, 404)


CHILD = r'''
import asyncio, json, statistics, sys, threading, time
arm, base = sys.argv[1], sys.argv[2]
if arm == "reference":
    from prime_sandboxes import APIClient, AsyncSandboxClient, SandboxClient
    from prime_evals import EvalsClient
    from prime_evals import APIClient as EvalsAPI
else:
    from prime_b200.platform.sandboxes import APIClient, AsyncSandboxClient, SandboxClient
    from prime_b200.platform.evals import EvalsClient
    from prime_b200.platform.evals import APIClient as EvalsAPI
out = {}
def med(f, reps):
    xs = []
    for _ in range(reps):
        t = time.perf_counter(); n = f(); xs.append(n / (time.perf_counter() - t))
    return round(statistics.median(xs), 1)
c = SandboxClient(APIClient(api_key="k"))
c.execute_command("s1", "true")  # auth + vm lookups cached from here on
def seq():
    for _ in range(400): c.execute_command("s1", "echo ok")
    return 400
out["sandbox_exec_sequential_calls_per_s"] = med(seq, 5)
def par():
    def work():
        for _ in range(100): c.execute_command("s1", "echo ok")
    ts = [threading.Thread(target=work) for _ in range(16)]
    [t.start() for t in ts]; [t.join() for t in ts]
    return 1600
out["sandbox_exec_16_threads_calls_per_s"] = med(par, 5)
async def fan():
    ac = AsyncSandboxClient(api_key="k")
    sem = asyncio.Semaphore(64)
    await ac.execute_command("s1", "true")
    async def one():
        async with sem:
            await ac.execute_command("s1", "echo ok")
    t = time.perf_counter()
    await asyncio.gather(*[one() for _ in range(2000)])
    dt = time.perf_counter() - t
    await ac.aclose()
    return 2000 / dt
out["async_fanout_2000_calls_sem64_per_s"] = round(statistics.median([asyncio.run(fan()) for _ in range(3)]), 1)
samples = [{"example_id": i, "task": "bench", "reward": 0.5, "prompt": [{"role": "user", "content": "x" * 600}],
            "completion": [{"role": "assistant", "content": "y" * 900}], "info": {"i": i}} for i in range(20000)]
ec = EvalsClient(EvalsAPI(api_key="k"))
def push():
    ec.push_samples("ev1", samples)
    return len(samples)
out["eval_push_20000_samples_per_s"] = med(push, 3)
print(json.dumps(out))
'''


def run_child(arm: str, base: str, home: str) -> dict:
    env = {**os.environ, "HOME": home, "PRIME_API_BASE_URL": base, "PRIME_BASE_URL": base, "PRIME_API_KEY": "k", "PRIME_DISABLE_VERSION_CHECK": "1",
           "PYTHONPATH": os.pathsep.join(REF_PATHS if arm == "reference" else [str(ROOT)])}  # fmt: skip
    r = subprocess.run([sys.executable, "-c", CHILD, arm, base], env=env, capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        return {"error": (r.stderr or r.stdout)[-1500:]}
    return json.loads(r.stdout.strip().splitlines()[-1])


def cold_start(arm: str, home: str, args: list[str]) -> float:
    code = "from prime_cli.main import run; run()" if arm == "reference" else "from prime_b200.platform.main import run; run()"
    env = {**os.environ, "HOME": home, "PRIME_DISABLE_VERSION_CHECK": "1", "PYTHONPATH": os.pathsep.join(REF_PATHS if arm == "reference" else [str(ROOT)])}
    xs = []
    for _ in range(7):
        t = time.perf_counter()
        r = subprocess.run([sys.executable, "-c", f"import sys; sys.argv = ['prime', *{args!r}]; {code}"], env=env, capture_output=True, text=True, timeout=120)
        xs.append(time.perf_counter() - t)
        if r.returncode != 0:
            raise SystemExit(f"{arm} {args}: rc {r.returncode}: {r.stderr[-800:]}")
    return round(statistics.median(xs) * 1e3, 1)


def main() -> None:
    ThreadingHTTPServer.request_queue_size = 512  # 64-way fan-outs connect at once; the default backlog of 5 resets connections
    srv = ThreadingHTTPServer(("127.0.0.1", 0), Handler)
    srv.daemon_threads = True
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    base = f"http://127.0.0.1:{srv.server_address[1]}"
    res: dict = {"method": __doc__.split("\n\n")[2].strip(), "python": sys.version.split()[0], "cpu_count": os.cpu_count()}
    with tempfile.TemporaryDirectory() as h1, tempfile.TemporaryDirectory() as h2:
        homes = {"reference": h1, "ours": h2}
        for arm in ("reference", "ours"):
            res[arm] = {"cli_help_ms": cold_start(arm, homes[arm], ["--help"]), "cli_sandbox_help_ms": cold_start(arm, homes[arm], ["sandbox", "--help"])}
        for arm in ("reference", "ours", "reference", "ours"):  # interleaved, second pass kept (warm page cache for both)
            res[arm].update(run_child(arm, base, homes[arm]))
    ratios = {}
    for k, v in res["ours"].items():
        rv = res["reference"].get(k)
        if isinstance(v, (int, float)) and isinstance(rv, (int, float)) and rv and v:
            ratios[k] = round(rv / v, 3) if k.endswith("_ms") else round(v / rv, 3)
    res["ours_vs_reference (>1 = this repo is faster)"] = ratios
    res["server_request_counts"] = Handler.counts
    print(json.dumps(res, indent=1))
    srv.shutdown()


if __name__ == "__main__":
    main()
