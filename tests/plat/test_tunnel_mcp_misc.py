"""Tunnel supervisor (real child process standing in for frpc), pinned-binary download, MCP tools, login challenge,
upgrade detection, SSE parsing (reference tests: packages/prime-tunnel/tests/*, packages/prime-mcp-server/tests/*,
packages/prime/tests/test_login.py, test_upgrade.py, test_inference.py)."""

import base64
import hashlib
import io
import os
import stat
import tarfile
from datetime import datetime, timedelta, timezone
from types import SimpleNamespace

import pytest
from cryptography.hazmat.primitives import hashes, serialization
from cryptography.hazmat.primitives.asymmetric import padding

from prime_b200.platform.api.inference import parse_sse_lines
from prime_b200.platform.commands import login as login_mod
from prime_b200.platform.commands import upgrade as up_mod
from prime_b200.platform.mcp import client as mcp_client
from prime_b200.platform.mcp.tools import pods as mcp_pods
from prime_b200.platform.tunnel import binary as bin_mod
from prime_b200.platform.tunnel import tunnel as tun_mod
from prime_b200.platform.tunnel.exceptions import BinaryDownloadError, TunnelConnectionError, TunnelTimeoutError
from prime_b200.platform.tunnel.models import TunnelInfo

INFO = TunnelInfo(tunnel_id="t-123", hostname="t-123.tunnel.example", url="https://t-123.tunnel.example", frp_token="tok-SECRET", binding_secret="bind",
                  server_host="frps.example", server_port=7000, expires_at=datetime.now(timezone.utc) + timedelta(hours=1))  # fmt: skip


class FakeTunnelAPI:
    def __init__(self):
        self.deleted = []

    async def create_tunnel(self, **kw):
        self.kw = kw
        return INFO

    async def delete_tunnel(self, tid):
        self.deleted.append(tid)
        return True

    async def close(self):
        pass


def fake_frpc(tmp_path, body: str):
    p = tmp_path / "frpc"
    p.write_text("#!/bin/bash\n" + body)
    p.chmod(p.stat().st_mode | stat.S_IXUSR)
    return p


@pytest.mark.anyio
async def test_tunnel_lifecycle_with_real_child(tmp_path, monkeypatch, isolated_home):
    frpc = fake_frpc(tmp_path, 'echo "login to server success"; echo "[proxy] start proxy success"; for i in $(seq 1 200); do echo "heartbeat $i"; sleep 0.01; done; sleep 60\n')
    monkeypatch.setattr(tun_mod, "get_frpc_path", lambda: frpc)
    api = FakeTunnelAPI()
    t = tun_mod.Tunnel(8000, name="demo", client=api, connection_timeout=10)
    url = await t.start()
    assert url == INFO.url and t.is_running and t.tunnel_id == "t-123" and api.kw == {"local_port": 8000, "name": "demo", "team_id": None}
    cfg = isolated_home / ".prime" / "tunnels" / "t-123.toml"
    assert stat.S_IMODE(cfg.stat().st_mode) == 0o600 and stat.S_IMODE(cfg.parent.stat().st_mode) == 0o700
    text = cfg.read_text()
    assert "tok-SECRET" in text and "localPort = 8000" in text and 'serverAddr = "frps.example"' in text
    import asyncio

    await asyncio.sleep(0.5)  # the drain threads keep consuming output into a bounded ring
    recent = t.recent_output
    assert 0 < len(recent) <= tun_mod.RING and any("heartbeat" in ln for ln in recent)
    with pytest.raises(Exception):
        await t.start()  # already started
    pid = t._process.pid
    await t.stop()
    assert not t.is_running and not cfg.exists() and api.deleted == ["t-123"]
    with pytest.raises(ProcessLookupError):
        os.kill(pid, 0)


@pytest.mark.anyio
async def test_tunnel_failure_modes_clean_up(tmp_path, monkeypatch, isolated_home):
    api = FakeTunnelAPI()
    monkeypatch.setattr(tun_mod, "get_frpc_path", lambda: fake_frpc(tmp_path, 'echo "login to server failed: authorization failed" >&2; exit 1\n'))
    with pytest.raises(TunnelConnectionError) as e:
        await tun_mod.Tunnel(8000, client=api, connection_timeout=5).start()
    assert "authorization failed" in str(e.value) and api.deleted == ["t-123"] and not list((isolated_home / ".prime" / "tunnels").glob("*.toml"))
    monkeypatch.setattr(tun_mod, "get_frpc_path", lambda: fake_frpc(tmp_path, 'echo "still connecting"; sleep 30\n'))
    t = tun_mod.Tunnel(8000, client=api, connection_timeout=0.5)
    with pytest.raises(TunnelTimeoutError) as e:
        await t.start()
    assert "still connecting" in str(e.value) and not t.is_running and api.deleted == ["t-123", "t-123"]


def _frp_tarball(payload=b"#!/bin/sh\necho frpc\n"):
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz") as t:
        ti = tarfile.TarInfo("frp_0.0.0_linux_amd64/frpc")
        ti.size, ti.mode = len(payload), 0o644
        t.addfile(ti, io.BytesIO(payload))
    return buf.getvalue()


def test_frpc_download_is_checksum_pinned_and_atomic(tmp_path, monkeypatch):
    blob = _frp_tarball()
    key = ("Linux", "x86_64")
    monkeypatch.setattr(bin_mod, "platform_key", lambda: key)
    dest = tmp_path / "bin" / "frpc"
    with pytest.raises(BinaryDownloadError, match="Checksum"):
        bin_mod.download_frpc(dest, fetch=lambda url: blob)  # not the pinned release → refused, nothing installed
    assert not dest.exists()
    monkeypatch.setitem(bin_mod.PINS, key, ("linux", "amd64", hashlib.sha256(blob).hexdigest()))
    urls = []
    bin_mod.download_frpc(dest, fetch=lambda url: (urls.append(url), blob)[1])
    assert dest.read_bytes().endswith(b"echo frpc\n") and os.access(dest, os.X_OK) and bin_mod.FRPC_VERSION in urls[0] and "linux_amd64" in urls[0]
    monkeypatch.setattr(bin_mod, "platform_key", lambda: ("Plan9", "mips"))
    with pytest.raises(BinaryDownloadError, match="Unsupported"):
        bin_mod.download_frpc(dest, fetch=lambda url: blob)
    # version stamp short-circuits the download
    cfg = SimpleNamespace(bin_dir=dest.parent)
    (dest.parent / ".frpc_version").write_text(bin_mod.FRPC_VERSION)
    monkeypatch.setattr(bin_mod, "download_frpc", lambda *a, **k: (_ for _ in ()).throw(AssertionError("must not download")))
    assert bin_mod.get_frpc_path(cfg) == dest


# ----------------------------------------------------------------------------------------------- MCP
class FakeAsyncAPI:
    def __init__(self, fail=False):
        self.calls, self.fail = [], fail

    async def _do(self, verb, endpoint, **kw):
        self.calls.append((verb, endpoint, kw))
        if self.fail:
            raise RuntimeError("boom")
        return {"ok": True, "endpoint": endpoint}

    async def get(self, endpoint, params=None):
        return await self._do("GET", endpoint, params=params)

    async def post(self, endpoint, json=None):
        return await self._do("POST", endpoint, json=json)

    async def patch(self, endpoint, json=None):
        return await self._do("PATCH", endpoint, json=json)

    async def delete(self, endpoint):
        return await self._do("DELETE", endpoint)


@pytest.mark.anyio
async def test_mcp_tools_never_raise(monkeypatch):
    api = FakeAsyncAPI()
    mcp_client.set_client(api)
    try:
        out = await mcp_pods.create_pod("cloud", "H100_80GB", "runpod", "dc-1", gpu_count=2, name="box", disk_size=100, env_vars={"A": "1"}, team_id="t1")
        assert out["ok"] and api.calls[0][0] == "POST"
        body = api.calls[0][2]["json"]
        assert body == {"pod": {"cloudId": "cloud", "gpuType": "H100_80GB", "gpuCount": 2, "socket": "PCIe", "image": "ubuntu_22_cuda_12", "dataCenterId": "dc-1",
                                "name": "box", "diskSize": 100, "envVars": [{"key": "A", "value": "1"}]}, "provider": {"type": "runpod"}, "team": {"teamId": "t1"}}  # fmt: skip
        assert (await mcp_pods.create_pod("c", "g", "p", "d", gpu_count=0)) == {"error": "gpu_count must be greater than 0"} and len(api.calls) == 1
        assert (await mcp_pods.create_pod("c", "g", "p", "d", disk_size=-5))["error"].startswith("disk_size")
        await mcp_pods.list_pods(offset=-3, limit=10)
        assert api.calls[-1][2]["params"] == {"offset": 0, "limit": 10}
        assert (await mcp_client.make_prime_request("PUT", "x")) == {"error": "Unsupported HTTP method: PUT"}
        mcp_client.set_client(FakeAsyncAPI(fail=True))
        assert (await mcp_pods.list_pods()) == {"error": "boom"}  # exceptions become data
    finally:
        mcp_client.set_client(None)


def test_mcp_server_registers_nine_tools():
    import asyncio

    from prime_b200.platform.mcp.server import mcp

    tools = asyncio.run(mcp.list_tools())
    assert sorted(t.name for t in tools) == sorted(["check_gpu_availability", "check_cluster_availability", "create_pod", "list_pods", "get_pods_history",
                                                    "get_pods_status", "get_pod_details", "delete_pod", "manage_ssh_keys"])  # fmt: skip


# ----------------------------------------------------------------------------------------------- login / upgrade / SSE
def test_login_challenge_roundtrip_decrypts_with_ephemeral_key():
    state = {"polls": 0}

    class Http:
        def post(self, url, json=None):
            state["pem"] = json["encryptionPublicKey"]
            return SimpleNamespace(status_code=200, json=lambda: {"challenge": "ABCD", "status_auth_token": "st"}, text="")

        def get(self, url, params=None, headers=None):
            state["polls"] += 1
            assert headers == {"Authorization": "Bearer st"} and params == {"challenge": "ABCD"}
            if state["polls"] < 3:
                return SimpleNamespace(status_code=200, json=lambda: {"result": None})
            pub = serialization.load_pem_public_key(state["pem"].encode())
            blob = pub.encrypt(b"pit_live_key", padding.OAEP(mgf=padding.MGF1(algorithm=hashes.SHA256()), algorithm=hashes.SHA256(), label=None))
            return SimpleNamespace(status_code=200, json=lambda: {"result": base64.b64encode(blob).decode()})

    seen = []
    key = login_mod.run_challenge("https://api", "https://app", http=Http(), announce=lambda url, code: seen.append((url, code)), sleep=lambda s: None)
    assert key == "pit_live_key" and seen == [("https://app/dashboard/tokens/challenge?code=ABCD", "ABCD")] and state["polls"] == 3

    class Expired(Http):
        def get(self, *a, **k):
            return SimpleNamespace(status_code=404, json=lambda: {})

    assert login_mod.run_challenge("https://api", "https://app", http=Expired(), announce=lambda *a: None, sleep=lambda s: None) is None


def test_upgrade_picks_the_installers_tool():
    assert up_mod.detect_install_method("/home/u/.local/share/uv/tools/prime/bin/python") == "uv_tool"
    assert up_mod.detect_install_method("/home/u/.local/pipx/venvs/prime/bin/python") == "pipx"
    assert up_mod.detect_install_method("/usr/bin/python3") == "pip"
    ran = []
    ok = up_mod.run_upgrade("pip", runner=lambda cmd, **kw: (ran.append(cmd), SimpleNamespace(returncode=0 if cmd[0] == "pip" else 1, stderr="no uv env"))[1],
                            which=lambda t: "/usr/bin/" + t)  # fmt: skip
    assert ok and [c[0] for c in ran] == ["uv", "pip"]  # falls through to the next recipe when the first fails
    assert not up_mod.run_upgrade("pipx", runner=lambda *a, **k: SimpleNamespace(returncode=0, stderr=""), which=lambda t: None)  # tool not installed
    # version/flag logic: --check only reports, --force upgrades even when current, otherwise nothing to do
    assert up_mod.decide("1.0.0", "1.1.0", check=True, force=False) == ("report", True)
    assert up_mod.decide("1.1.0", "1.1.0", check=False, force=True) == ("upgrade", False)
    assert up_mod.decide("1.1.0", "1.1.0", check=False, force=False) == ("current", False)
    assert up_mod.decide("1.0.0", "1.1.0", check=False, force=False) == ("upgrade", True)


def test_sse_parser():
    lines = ["", 'data: {"id": 1}', ": keep-alive comment is not json", '{"id": 2}', "data: not-json", "data: [DONE]", 'data: {"id": 3}']
    assert [c["id"] for c in parse_sse_lines(lines)] == [1, 2]
