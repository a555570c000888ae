"""Tunnel lifecycle: register → write 0600 frpc config → spawn frpc → wait for "start proxy success" →
drain pipes in daemon threads (ring buffer of the last 50 lines) → stop / signal-safe sync_stop / cleanup
(reference: packages/prime-tunnel/src/prime_tunnel/tunnel.py:59-465)."""

from __future__ import annotations

import asyncio
import collections
import os
import re
import subprocess
import threading
import time
from pathlib import Path

import httpx

from .binary import get_frpc_path
from .client import TunnelClient
from .exceptions import TunnelConnectionError, TunnelError, TunnelTimeoutError
from .models import TunnelInfo

_FRPC_LOG = re.compile(r"\d{4}-\d{2}-\d{2}\s\d{2}:\d{2}:\d{2}\.\d{3}\s\[([EWIDT])\]\s\[.*?\]\s(?:\[.*?\]\s)*(.+)")
_ANSI = re.compile(r"\x1b\[[0-9;]*m")
SUCCESS_MARK = "start proxy success"
FAILURE_MARKS = ("login to the server failed", "connect to server error")
RING = 50


def frpc_failure(lines: list[str], tunnel_id: str | None = None, return_code: int | None = None) -> TunnelConnectionError:
    """Pick the last error/warning-level frpc log message; fall back to the raw output."""
    problems = [m.group(2) for m in (_FRPC_LOG.match(_ANSI.sub("", ln)) for ln in lines) if m and m.group(1) in "EW"]
    if problems:
        return TunnelConnectionError(problems[-1], tunnel_id=tunnel_id)
    text = "\n".join(lines) if lines else "(no output captured)"
    code = f" (exit code {return_code})" if return_code is not None else ""
    return TunnelConnectionError(f"frpc process failed{code}: {text}", tunnel_id=tunnel_id)


def render_frpc_config(info: TunnelInfo, local_addr: str, local_port: int, log_level: str) -> str:
    t = info.tunnel_id
    return "\n".join([
        "# Prime tunnel frpc configuration", f"# Tunnel ID: {t}",
        f'serverAddr = "{info.server_host}"', f"serverPort = {info.server_port}",
        f'user = "{t}"', 'auth.method = "token"', f'auth.token = "{info.frp_token}"',
        f'metadatas.binding_secret = "{info.binding_secret}"',
        "transport.tcpMux = true", "transport.tcpMuxKeepaliveInterval = 30", "transport.poolCount = 10",
        "transport.dialServerKeepalive = 60",
        'log.to = "console"', f'log.level = "{log_level}"',  # console: readiness is detected on stdout
        "[[proxies]]", f'name = "{t}"', 'type = "http"', f'localIP = "{local_addr}"', f"localPort = {local_port}",
        f'subdomain = "{t}"', "",
    ])  # fmt: skip


class Tunnel:
    def __init__(self, local_port: int, local_addr: str = "127.0.0.1", name: str | None = None, connection_timeout: float = 30.0,
                 log_level: str = "info", team_id: str | None = None, client: TunnelClient | None = None):  # fmt: skip
        self.local_port, self.local_addr, self.name, self.team_id = local_port, local_addr, name, team_id
        self.connection_timeout, self.log_level = connection_timeout, log_level
        self._client = client or TunnelClient()
        self._process: subprocess.Popen | None = None
        self._info: TunnelInfo | None = None
        self._config_file: Path | None = None
        self._started = False
        self._lines: collections.deque[str] = collections.deque(maxlen=RING)
        self._lock = threading.Lock()
        self._drainers: list[threading.Thread] = []

    # ---- facts
    @property
    def _tunnel_info(self) -> TunnelInfo | None:  # the reference's name of the same private slot (its tests and subclasses reach for it)
        return self._info

    @_tunnel_info.setter
    def _tunnel_info(self, value: TunnelInfo | None) -> None:
        self._info = value

    @property
    def tunnel_id(self) -> str | None:
        return self._info.tunnel_id if self._info else None

    @property
    def url(self) -> str | None:
        return self._info.url if self._info else None

    @property
    def hostname(self) -> str | None:
        return self._info.hostname if self._info else None

    @property
    def is_running(self) -> bool:
        return self._process is not None and self._process.poll() is None

    @property
    def recent_output(self) -> list[str]:
        if not self.is_running:
            for t in self._drainers:
                t.join(timeout=2.0)
        with self._lock:
            return list(self._lines)

    # ---- lifecycle
    async def start(self) -> str:
        if self._started:
            raise TunnelError("Tunnel is already started")
        frpc = await asyncio.to_thread(get_frpc_path)
        try:
            try:
                self._info = await self._client.create_tunnel(local_port=self.local_port, name=self.name, team_id=self.team_id)
            except (TunnelError, asyncio.CancelledError):
                raise
            except Exception as e:
                raise TunnelError(f"Failed to register tunnel: {e}") from e
            try:
                self._config_file = self._write_frpc_config()
            except OSError as e:
                raise TunnelError(f"Failed to write frpc config: {e}") from e
            try:
                self._process = subprocess.Popen([str(frpc), "-c", str(self._config_file)], stdout=subprocess.PIPE,
                                                 stderr=subprocess.PIPE, text=True)  # fmt: skip
            except OSError as e:
                raise TunnelConnectionError(f"Failed to start frpc: {e}") from e
            await self._wait_for_connection()
            self._start_pipe_drain()
        except BaseException:
            await self._cleanup()
            raise
        self._started = True
        return self.url  # type: ignore[return-value]

    async def stop(self) -> None:
        if self._started:
            await self._cleanup()
            self._started = False

    def _kill_process(self) -> None:
        proc, self._process = self._process, None
        if proc is None:
            return
        try:
            proc.terminate()
            try:
                proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                proc.kill()
                proc.wait(timeout=2)
        except Exception:
            pass

    def _drop_config(self) -> None:
        cfg, self._config_file = self._config_file, None
        try:
            if cfg is not None:
                cfg.unlink(missing_ok=True)
        except Exception:
            pass

    def sync_stop(self) -> None:
        """Blocking teardown usable from a signal handler (no event loop needed)."""
        if not self._started:
            return
        self._kill_process()
        info, self._info = self._info, None
        if info is not None:
            try:
                httpx.delete(f"{self._client.base_url}/api/v1/tunnel/{info.tunnel_id}", headers=self._client._headers, timeout=5.0)
            except Exception:
                pass
        self._drop_config()
        self._started = False

    async def _cleanup(self) -> None:
        self._kill_process()  # EOF on the pipes ends the drain threads
        info, self._info = self._info, None
        if info is not None:
            try:
                await self._client.delete_tunnel(info.tunnel_id)
            except Exception:
                pass
        self._drop_config()
        try:
            await self._client.close()
        except Exception:
            pass

    # ---- internals
    def _write_frpc_config(self) -> Path:
        if self._info is None:
            raise TunnelError("Tunnel not registered")
        d = Path.home() / ".prime" / "tunnels"
        d.mkdir(parents=True, exist_ok=True)
        d.chmod(0o700)
        path = d / f"{self._info.tunnel_id}.toml"
        fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)  # token inside: owner-only from birth
        try:
            os.write(fd, render_frpc_config(self._info, self.local_addr, self.local_port, self.log_level).encode())
        finally:
            os.close(fd)
        return path

    def _record(self, line: str) -> None:
        with self._lock:
            self._lines.append(line)

    def _poll_pipes(self) -> str | None:
        """Non-blocking read of whatever frpc has printed; returns "ok" / "fail" when a marker shows up."""
        assert self._process is not None
        verdict = None
        for pipe in (self._process.stdout, self._process.stderr):
            if pipe is None:
                continue
            fd = pipe.fileno()
            was_blocking = os.get_blocking(fd)
            os.set_blocking(fd, False)
            try:
                while True:
                    try:
                        line = pipe.readline()
                    except (BlockingIOError, OSError):
                        break
                    if not line:
                        break
                    line = line.strip()
                    if not line:
                        continue
                    self._record(line)
                    low = line.lower()
                    if SUCCESS_MARK in low:
                        verdict = verdict or "ok"
                    elif any(m in low for m in FAILURE_MARKS):
                        verdict = "fail"
            finally:
                try:
                    os.set_blocking(fd, was_blocking)
                except (OSError, ValueError):
                    pass
        return verdict

    async def _wait_for_connection(self) -> None:
        deadline = time.monotonic() + self.connection_timeout
        while time.monotonic() < deadline:
            if self._process is None:
                raise TunnelConnectionError("frpc process not running")
            rc = self._process.poll()
            if rc is not None:
                for pipe in (self._process.stdout, self._process.stderr):
                    if pipe:
                        for ln in pipe.readlines():
                            if ln.strip():
                                self._record(ln.strip())
                raise frpc_failure(list(self._lines), self.tunnel_id, rc)
            verdict = self._poll_pipes()
            if verdict == "ok":
                return
            if verdict == "fail":
                raise frpc_failure(list(self._lines), self.tunnel_id)
            await asyncio.sleep(0.1)
        text = "\n".join(self._lines) if self._lines else "(no output captured)"
        raise TunnelTimeoutError(f"Tunnel connection timed out after {self.connection_timeout}s\n--- frpc output ---\n{text}\n-------------------")

    def _start_pipe_drain(self) -> None:
        """frpc keeps logging (reconnects …); undrained pipes would eventually block it."""
        if self._process is None:
            return

        def drain(pipe) -> None:
            try:
                for line in pipe:
                    line = line.rstrip("\n")
                    if line:
                        self._record(line)
            except (OSError, ValueError):
                pass

        for pipe in (self._process.stdout, self._process.stderr):
            if pipe is not None:
                t = threading.Thread(target=drain, args=(pipe,), daemon=True)
                t.start()
                self._drainers.append(t)

    async def __aenter__(self):
        await self.start()
        return self

    async def __aexit__(self, *exc):
        await self.stop()
