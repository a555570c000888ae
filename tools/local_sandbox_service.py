#!/usr/bin/env python
"""A local, stateful stand-in for the sandbox control plane AND its gateway that really executes: every sandbox is a scratch directory,
``execute_command`` runs the command with ``bash -c`` inside it, uploads / downloads / ``read-file`` move real bytes, an exposed port is
the local port the sandbox's own server listens on.  For running SDK programs end to end with no network — the reference's example
scripts in ``tools/run_reference_examples.py``, or your own while offline:

    python tools/local_sandbox_service.py --port 8765 &
    PRIME_API_BASE_URL=http://127.0.0.1:8765 PRIME_API_KEY=local python examples/sandbox_quickstart.py

Paths: a sandbox has no root file system of its own, so the literal prefixes ``/sandbox-workspace``, ``/workspace`` and ``/tmp`` in commands
and file paths are rewritten to directories under the sandbox's scratch root (the default working directory is its ``/sandbox-workspace``).
NOT a security boundary: commands run as this user on this machine.  The wire shapes are the ones ``tools/wire_diff.py`` verifies against
the reference client (reference: packages/prime-sandboxes/src/prime_sandboxes/sandbox.py — REST exec / upload / download / read-file,
models.py:30-63 for the sandbox record).
"""

from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import subprocess
import tempfile
import threading
import time
import uuid
from datetime import datetime, timedelta, timezone
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from pathlib import Path
from urllib.parse import parse_qs, urlparse

VIRTUAL_ROOTS = ("/sandbox-workspace", "/workspace", "/tmp")


def multipart_file(content_type: str, raw: bytes, field: str) -> bytes:
    """The bytes of one named part of a multipart/form-data body (the `email` parser; `cgi` is gone in Python 3.13)."""
    from email.parser import BytesParser
    from email.policy import HTTP

    msg = BytesParser(policy=HTTP).parsebytes(b"Content-Type: " + content_type.encode() + b"\r\nMIME-Version: 1.0\r\n\r\n" + raw)
    for part in msg.iter_parts() if msg.is_multipart() else ():
        if part.get_param("name", header="content-disposition") == field:
            return part.get_payload(decode=True) or b""
    return b""


def now() -> str:
    return datetime.now(timezone.utc).isoformat()


class Sandbox:
    def __init__(self, body: dict, root: Path):
        self.id = "sbx-" + uuid.uuid4().hex[:12]
        self.root = root / self.id
        for v in VIRTUAL_ROOTS:
            (self.root / v.lstrip("/")).mkdir(parents=True, exist_ok=True)
        self.body, self.status, self.polls = body, "PENDING", 0
        self.created = now()
        self.started: str | None = None
        self.terminated: str | None = None
        self.exposures: dict[str, dict] = {}
        self.log: list[str] = [f"[{self.created}] container created from {body.get('docker_image')}"]
        self.children: list[subprocess.Popen] = []

    # ---- path / command virtualisation
    def host_path(self, path: str) -> Path:
        p = path if path.startswith("/") else f"/sandbox-workspace/{path}"
        return self.root / os.path.normpath(p).lstrip("/")

    def rewrite(self, text: str) -> str:
        """One pass (the scratch root itself may live under /tmp: a second pass would rewrite the rewritten)."""
        alt = "|".join(re.escape(v) for v in VIRTUAL_ROOTS)
        return re.sub(rf"(?<![\w.-])({alt})(?=$|[/\s\"';|&)>])", lambda m: str(self.root / m.group(1).lstrip("/")), text)

    def record(self) -> dict:
        b = self.body
        return {"id": self.id, "name": b.get("name", self.id), "dockerImage": b.get("docker_image", "python:3.11-slim"), "startCommand": b.get("start_command"),
                "cpuCores": b.get("cpu_cores", 1), "memoryGB": b.get("memory_gb", 2), "diskSizeGB": b.get("disk_size_gb", 10),
                "diskMountPath": "/sandbox-workspace", "gpuCount": b.get("gpu_count", 0), "gpuType": b.get("gpu_type"), "vm": bool(b.get("vm", False)),
                "networkAccess": b.get("network_access", True), "status": self.status, "timeoutMinutes": b.get("timeout_minutes", 60),
                "environmentVars": b.get("environment_vars"), "labels": b.get("labels") or [], "createdAt": self.created, "updatedAt": now(),
                "startedAt": self.started, "terminatedAt": self.terminated, "userId": "local-user", "teamId": b.get("team_id"),
                "kubernetesJobId": self.id}  # fmt: skip

    def tick(self) -> None:
        """PENDING on creation, RUNNING from the second look on — enough for ``wait_for_creation`` loops to do a real poll."""
        self.polls += 1
        if self.status == "PENDING" and self.polls >= 2:
            self.status, self.started = "RUNNING", now()
            self.log.append(f"[{self.started}] started: {self.body.get('start_command') or 'tail -f /dev/null'}")

    def run(self, command: str, working_dir: str | None, env: dict | None, timeout: float) -> tuple[int, dict]:
        if self.status != "RUNNING":
            self.tick()
            self.tick()
        cwd = self.host_path(working_dir) if working_dir else self.root / "sandbox-workspace"
        if not cwd.is_dir():
            return 200, {"stdout": "", "stderr": f"bash: cd: {working_dir}: No such file or directory\n", "exit_code": 1}
        full_env = {**os.environ, "HOME": str(self.root / "sandbox-workspace"), "SANDBOX_ID": self.id, "SANDBOX_NAME": str(self.body.get("name", "")),
                    **{k: str(v) for k, v in (self.body.get("environment_vars") or {}).items()},
                    **{k: str(v) for k, v in (self.body.get("secrets") or {}).items()},  # injected like variables, never echoed in the record
                    **{k: str(v) for k, v in (env or {}).items()}}  # fmt: skip
        for k in [k for k in full_env if k.startswith("PRIME_")]:
            del full_env[k]
        self.log.append(f"[{now()}] exec: {command[:200]}")
        proc = subprocess.Popen(["bash", "-c", self.rewrite(command)], cwd=cwd, env=full_env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                start_new_session=True)  # fmt: skip
        self.children.append(proc)
        try:
            out, err = proc.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            try:
                os.killpg(proc.pid, 9)
            except OSError:
                pass
            proc.communicate()
            return 408, {"error": "command timed out"}
        return 200, {"stdout": out.decode(errors="replace"), "stderr": err.decode(errors="replace"), "exit_code": proc.returncode}

    def destroy(self) -> None:
        self.status, self.terminated = "TERMINATED", now()
        for p in self.children:  # the exact process groups this sandbox started (a group outlives its leader while `nohup … &` children run)
            try:
                os.killpg(p.pid, 9)
            except OSError:
                pass
        shutil.rmtree(self.root, ignore_errors=True)


class Service:
    def __init__(self, root: Path | None = None):
        self.root = root or Path(tempfile.mkdtemp(prefix="prime_local_sandboxes_"))
        self.sandboxes: dict[str, Sandbox] = {}
        self.lock = threading.Lock()
        self.requests: list[tuple[str, str]] = []

    def get(self, sid: str) -> Sandbox | None:
        return self.sandboxes.get(sid)

    def close(self) -> None:
        for s in list(self.sandboxes.values()):
            if s.status != "TERMINATED":
                s.destroy()
        shutil.rmtree(self.root, ignore_errors=True)


def make_handler(svc: Service):
    class Handler(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def log_message(self, *a) -> None:  # quiet
            pass

        # ---- plumbing
        def _send(self, code: int, body) -> None:
            raw = body if isinstance(body, bytes) else json.dumps(body).encode()
            self.send_response(code)
            self.send_header("Content-Type", "application/octet-stream" if isinstance(body, bytes) else "application/json")
            self.send_header("Content-Length", str(len(raw)))
            self.end_headers()
            self.wfile.write(raw)

        def _body(self) -> bytes:
            n = int(self.headers.get("Content-Length") or 0)
            return self.rfile.read(n) if n else b""

        def _json(self) -> dict:
            raw = self._body()
            try:
                return json.loads(raw) if raw else {}
            except ValueError:
                return {}

        def _route(self, method: str) -> None:
            u = urlparse(self.path)
            q = {k: v[0] for k, v in parse_qs(u.query).items()}
            svc.requests.append((method, u.path))
            if not (self.headers.get("Authorization") or "").startswith("Bearer "):
                self._body()
                return self._send(401, {"detail": "missing bearer token"})
            try:
                if u.path.startswith("/gw/"):
                    return self._gateway(method, u.path, q)
                return self._control(method, u.path.removeprefix("/api/v1"), q)
            except BrokenPipeError:
                pass
            except Exception as e:  # noqa: BLE001
                self._send(500, {"detail": f"{type(e).__name__}: {e}"})

        do_GET = lambda self: self._route("GET")  # noqa: E731
        do_POST = lambda self: self._route("POST")  # noqa: E731
        do_DELETE = lambda self: self._route("DELETE")  # noqa: E731
        do_PATCH = lambda self: self._route("PATCH")  # noqa: E731

        # ---- control plane
        def _control(self, method: str, path: str, q: dict) -> None:
            host = self.headers.get("Host")
            m = re.fullmatch(r"/sandbox/?", path)
            if m and method == "POST":
                body = self._json()
                with svc.lock:
                    s = Sandbox(body, svc.root)
                    svc.sandboxes[s.id] = s
                return self._send(200, s.record())
            if m and method == "GET":
                labels = parse_qs(urlparse(self.path).query).get("labels", [])
                for s in svc.sandboxes.values():  # a listing is a look too (bulk_wait_for_creation polls the list, not the records)
                    s.tick()
                rows = [s for s in svc.sandboxes.values()
                        if (q.get("status") is None or s.status == q["status"])
                        and (q.get("is_active", "").lower() != "true" or s.status != "TERMINATED")
                        and all(lb in (s.body.get("labels") or []) for lb in labels)]  # fmt: skip
                page, per = int(q.get("page", 1)), int(q.get("per_page", 50))
                chunk = rows[(page - 1) * per : page * per]
                return self._send(200, {"sandboxes": [s.record() for s in chunk], "total": len(rows), "page": page, "perPage": per,
                                        "hasNext": page * per < len(rows)})  # fmt: skip
            if m and method == "DELETE":
                body = self._json()
                ids = body.get("sandbox_ids") or [s.id for s in svc.sandboxes.values() if set(body.get("labels") or []) & set(s.body.get("labels") or [])]
                ok, bad = [], []
                for i in ids:
                    s = svc.get(i)
                    if s is None:
                        bad.append({"sandbox_id": i, "error": "not found"})
                    else:
                        s.destroy()
                        ok.append(i)
                return self._send(200, {"succeeded": ok, "failed": bad, "message": f"Deleted {len(ok)} sandbox(es)"})
            if path == "/sandbox/expose/all" and method == "GET":
                self._body()
                return self._send(200, {"exposures": [e for s in svc.sandboxes.values() for e in s.exposures.values()]})
            m = re.fullmatch(r"/sandbox/([^/]+)(/.*)?", path)
            if not m:
                self._body()
                return self._send(404, {"detail": f"no such route: {method} {path}"})
            s, rest = svc.get(m.group(1)), m.group(2) or ""
            if s is None:
                self._body()
                return self._send(404, {"detail": f"Sandbox {m.group(1)} not found"})
            if rest == "" and method == "GET":
                s.tick()
                return self._send(200, s.record())
            if rest == "" and method == "DELETE":
                self._body()
                s.destroy()
                return self._send(200, {"status": "deleted", "id": s.id})
            if rest == "/logs":
                return self._send(200, {"logs": "\n".join(s.log) + "\n"})
            if rest == "/error-context":
                return self._send(200, {"status": s.status, "error_type": None, "error_message": None})
            if rest == "/auth" and method == "POST":
                self._body()
                if s.status == "TERMINATED":
                    return self._send(409, {"detail": "sandbox is terminated"})
                s.tick()
                s.tick()
                exp = (datetime.now(timezone.utc) + timedelta(hours=1)).isoformat()
                return self._send(200, {"gateway_url": f"http://{host}/gw", "user_ns": "local", "job_id": s.id, "token": "local-" + s.id, "expires_at": exp,
                                        "is_vm": bool(s.body.get("vm", False))})  # fmt: skip
            if rest == "/expose" and method == "POST":
                body = self._json()
                port, proto = int(body["port"]), (body.get("protocol") or "HTTP").upper()
                e = {"exposure_id": "exp-" + uuid.uuid4().hex[:8], "sandbox_id": s.id, "port": port, "name": body.get("name"),
                     "url": f"http://127.0.0.1:{port}", "tls_socket": f"127.0.0.1:{port}", "protocol": proto, "external_port": port,
                     "external_endpoint": f"127.0.0.1:{port}", "created_at": now()}  # fmt: skip
                s.exposures[e["exposure_id"]] = e
                return self._send(200, e)
            if rest == "/expose" and method == "GET":
                return self._send(200, {"exposures": list(s.exposures.values())})
            m2 = re.fullmatch(r"/expose/([^/]+)", rest)
            if m2 and method == "DELETE":
                self._body()
                return self._send(200 if s.exposures.pop(m2.group(1), None) else 404, {})
            self._body()
            return self._send(404, {"detail": f"no such route: {method} {path}"})

        # ---- gateway
        def _gateway(self, method: str, path: str, q: dict) -> None:
            m = re.fullmatch(r"/gw/([^/]+)/([^/]+)/([a-z-]+)", path)
            s = svc.get(m.group(2)) if m else None
            if s is None or s.status == "TERMINATED":
                self._body()
                return self._send(502, {"error": "sandbox_not_found"})
            op = m.group(3)
            if op == "exec" and method == "POST":
                b = self._json()
                code, out = s.run(b.get("command", ""), b.get("working_dir"), b.get("env"), float(b.get("timeout") or 300))
                return self._send(code, out)
            if op == "upload" and method == "POST":
                ctype = self.headers.get("Content-Type", "")
                raw = self._body()
                data = multipart_file(ctype, raw, "file")
                dest = s.host_path(q.get("path", "upload.bin"))
                dest.parent.mkdir(parents=True, exist_ok=True)
                dest.write_bytes(data)
                return self._send(200, {"success": True, "path": q.get("path"), "size": len(data), "timestamp": now()})
            if op in ("download", "read-file") and method == "GET":
                src = s.host_path(q.get("path", ""))
                if not src.is_file():
                    return self._send(404, {"detail": f"file not found: {q.get('path')}"})
                data = src.read_bytes()
                if op == "download":
                    return self._send(200, data)
                return self._send(200, {"content": data.decode(errors="replace"), "size": len(data)})
            self._body()
            return self._send(404, {"detail": f"no such gateway route: {method} {path}"})

    return Handler


def serve(port: int = 0) -> tuple[ThreadingHTTPServer, Service]:
    svc = Service()
    ThreadingHTTPServer.request_queue_size = 512
    httpd = ThreadingHTTPServer(("127.0.0.1", port), make_handler(svc))
    httpd.daemon_threads = True
    threading.Thread(target=httpd.serve_forever, daemon=True).start()
    return httpd, svc


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--port", type=int, default=8765)
    a = ap.parse_args()
    httpd, svc = serve(a.port)
    print(f"local sandbox service on http://127.0.0.1:{httpd.server_address[1]}  (scratch: {svc.root})\n"
          f"  export PRIME_API_BASE_URL=http://127.0.0.1:{httpd.server_address[1]} PRIME_API_KEY=local", flush=True)  # fmt: skip
    try:
        while True:
            time.sleep(3600)
    except KeyboardInterrupt:
        pass
    finally:
        httpd.shutdown()
        svc.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
