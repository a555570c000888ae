"""frpc binary manager: pinned release, SHA-256 verified, atomically installed under ``~/.prime/bin``
(reference: packages/prime-tunnel/src/prime_tunnel/binary.py:15-155).  The pin (version, URLs, digests) is
data that must match upstream frp release artefacts; the installer logic is ours."""

from __future__ import annotations

import hashlib
import os
import platform
import stat
import tarfile
import tempfile
from pathlib import Path

import httpx

from ..core.config import Config
from .exceptions import BinaryDownloadError

FRPC_VERSION = "0.66.0"
_RELEASE = "https://github.com/fatedier/frp/releases/download/v{v}/frp_{v}_{os}_{arch}.tar.gz"
# (platform.system(), normalised machine) → (release os, release arch, sha256 of the tarball)
PINS: dict[tuple[str, str], tuple[str, str, str]] = {
    ("Darwin", "arm64"): ("darwin", "arm64", "eb24c3c172a20056d83379496500b92600a992f68e8ae2e27d128ce1f36d7a92"),
    ("Darwin", "x86_64"): ("darwin", "amd64", "9558d55a9d8bc40e22018379ea645251f803f9e2d69e7a7a2fd1588f98f8ef43"),
    ("Linux", "x86_64"): ("linux", "amd64", "317a17a7adac2e6bed2d7a83dc077da91ced0d110e1636373ece8ae5ac8b578b"),
    ("Linux", "aarch64"): ("linux", "arm64", "196ddaa51b716c2e99aeb2916b0a2bf55bb317494c4acdcefab36c383de950ba"),
}


def platform_key() -> tuple[str, str]:
    system, machine = platform.system(), platform.machine()
    if machine in ("AMD64", "x86_64"):
        machine = "x86_64"
    elif machine in ("arm64", "aarch64"):
        machine = "arm64" if system == "Darwin" else "aarch64"
    return system, machine


def release_url(key: tuple[str, str]) -> str:
    os_name, arch, _ = PINS[key]
    return _RELEASE.format(v=FRPC_VERSION, os=os_name, arch=arch)


def sha256_file(path: Path) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 16), b""):
            h.update(chunk)
    return h.hexdigest()


def atomic_write(dest: Path, data: bytes | None = None, src: Path | None = None, mode: int | None = None) -> None:
    """temp-in-same-dir + rename, so concurrent installers can never expose a partial file."""
    dest.parent.mkdir(parents=True, exist_ok=True)
    tmp = dest.parent / f".{dest.name}.{os.getpid()}.tmp"
    try:
        tmp.write_bytes(data if data is not None else Path(src).read_bytes())
        if mode is not None:
            tmp.chmod(mode)
        os.replace(tmp, dest)
    finally:
        tmp.unlink(missing_ok=True)


def extract_frpc(archive: Path, into: Path) -> Path:
    try:
        with tarfile.open(archive, "r:gz") as tar:
            member = next((m for m in tar.getmembers() if m.isfile() and (m.name == "frpc" or m.name.endswith("/frpc"))), None)
            if member is None:
                raise BinaryDownloadError("frpc binary not found in archive")
            member.name = "frpc"  # flatten: never trust archive paths
            tar.extract(member, into)
    except tarfile.TarError as e:
        raise BinaryDownloadError(f"Failed to extract frpc: {e}") from e
    return into / "frpc"


def download_frpc(dest: Path, fetch=None) -> None:
    key = platform_key()
    if key not in PINS:
        raise BinaryDownloadError(f"Unsupported platform: {key[0]} {key[1]}")
    expected = PINS[key][2]
    with tempfile.TemporaryDirectory() as tmp:
        archive = Path(tmp) / "frp.tar.gz"
        try:
            if fetch is not None:
                archive.write_bytes(fetch(release_url(key)))
            else:
                with httpx.stream("GET", release_url(key), follow_redirects=True, timeout=120.0) as r:
                    r.raise_for_status()
                    with open(archive, "wb") as f:
                        for chunk in r.iter_bytes(1 << 16):
                            f.write(chunk)
        except httpx.HTTPError as e:
            raise BinaryDownloadError(f"Failed to download frpc: {e}") from e
        got = sha256_file(archive)
        if got != expected:
            raise BinaryDownloadError(f"Checksum verification failed: expected {expected}, got {got}")
        binary = extract_frpc(archive, Path(tmp))
        exec_mode = binary.stat().st_mode | stat.S_IXUSR | stat.S_IXGRP | stat.S_IXOTH
        atomic_write(dest, src=binary, mode=exec_mode)


def get_frpc_path(config: Config | None = None) -> Path:
    bin_dir = (config or Config(writable=False)).bin_dir
    frpc, stamp = bin_dir / "frpc", bin_dir / ".frpc_version"
    if frpc.exists() and stamp.exists() and stamp.read_text().strip() == FRPC_VERSION:
        return frpc
    download_frpc(frpc)
    atomic_write(stamp, data=FRPC_VERSION.encode())
    return frpc
