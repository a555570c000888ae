"""File round-trips with checksums: small text, binary, and (with --stress) 10/20/30 MB payloads timed both ways."""

import argparse
import hashlib
import os
import tempfile
import time
from pathlib import Path

from prime_b200.platform.sandboxes import APIClient, CreateSandboxRequest, SandboxClient, SandboxFileNotFoundError

ap = argparse.ArgumentParser()
ap.add_argument("--stress", action="store_true")
a = ap.parse_args()
client = SandboxClient(APIClient())
sb = client.create(CreateSandboxRequest(name="files-demo", docker_image="python:3.11-slim", timeout_minutes=15, disk_size_gb=5))
try:
    client.wait_for_creation(sb.id)
    with tempfile.TemporaryDirectory() as td:
        sizes = [("hello.txt", b"hello\n"), ("blob.bin", os.urandom(256 * 1024))]
        if a.stress:
            sizes += [(f"big_{mb}mb.bin", os.urandom(mb << 20)) for mb in (10, 20, 30)]
        for name, payload in sizes:
            src, back = Path(td) / name, Path(td) / f"back_{name}"
            src.write_bytes(payload)
            t0 = time.perf_counter()
            client.upload_file(sb.id, f"/tmp/{name}", str(src), timeout=300)
            t1 = time.perf_counter()
            remote = client.execute_command(sb.id, f"sha256sum /tmp/{name} | cut -d' ' -f1").stdout.strip()
            client.download_file(sb.id, f"/tmp/{name}", str(back), timeout=300)
            t2 = time.perf_counter()
            local = hashlib.sha256(payload).hexdigest()
            ok = remote == local == hashlib.sha256(back.read_bytes()).hexdigest()
            mb = len(payload) / 2**20
            print(f"{name:14s} {mb:7.2f} MiB  up {mb / (t1 - t0):6.1f} MiB/s  down {mb / (t2 - t1):6.1f} MiB/s  checksum {'ok' if ok else 'MISMATCH'}")
    client.upload_bytes(sb.id, "/tmp/inline.json", b'{"inline": true}', "inline.json")
    print("read_file:", client.read_file(sb.id, "/tmp/inline.json").content)
    try:
        client.download_file(sb.id, "/tmp/does-not-exist", "/tmp/nope")
    except SandboxFileNotFoundError as e:
        print("missing file →", type(e).__name__)
finally:
    client.delete(sb.id)
