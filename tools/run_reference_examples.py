#!/usr/bin/env python
"""Run the reference's example programs (``/root/reference/examples/*.py``), unmodified, on this package — the "switch" test a user
would do: same script, ``prime_sandboxes`` now resolving to ``prime_b200.platform.sandboxes`` — and on the reference itself as the control.

    python tools/run_reference_examples.py                 # → profiles/reference_examples.json
    python tools/run_reference_examples.py --only sandbox_demo -v

Both arms talk to ``tools/local_sandbox_service.py`` (a fresh instance per run): sandboxes are scratch directories, commands really run,
files really move, exposed ports are really reachable — so an example's own checks (file sizes, checksums, HTTP fetch of an exposed port)
are meaningful.  Compared per example: exit code, the (method, route) sequence the service saw with ids removed, and stdout with ids,
paths, timings and sizes-per-second masked.  ``sandbox_tailscale.py`` needs a Tailscale key and daemon and is skipped in both arms.
"""

from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
REFERENCE = Path(os.environ.get("PRIME_REFERENCE_ROOT", "/root/reference"))
SKIP = {"sandbox_tailscale": "needs a Tailscale auth key and tailscaled inside the sandbox"}
ARGS = {"sandbox_file_handling_stress_test": ["sequential"]}
SDK_DIRS = ("prime", "prime-evals", "prime-sandboxes", "prime-tunnel", "prime-mcp-server")

_BOOT = """
import runpy, sys
impl, script = sys.argv[1], sys.argv[2]
sys.argv = [script] + sys.argv[3:]
if impl == "ours":
    sys.path.insert(0, {repo!r})
    from prime_b200 import compat
    compat.install()
else:
    for d in {sdk!r}:
        sys.path.insert(0, {ref!r} + "/packages/" + d + "/src")
runpy.run_path(script, run_name="__main__")
"""


def mask(text: str) -> str:
    text = re.sub(r"sbx-[0-9a-f]{12}", "sbx-ID", text)
    text = re.sub(r"exp-[0-9a-f]{8}", "exp-ID", text)
    text = re.sub(r"/tmp/prime_local_sandboxes_\w+", "/ROOT", text)
    text = re.sub(r"/tmp/[\w./-]+", "/tmp/PATH", text)
    text = re.sub(r"127\.0\.0\.1:\d+", "HOST:PORT", text)
    text = re.sub(r"\b[0-9a-f]{8}\b", "HEX8", text)
    text = re.sub(r"\d{4}-\d\d-\d\d[T ]\d\d:\d\d:\d\d[.\d]*(\+00:00|Z)?", "TIME", text)
    text = re.sub(r"\d+(\.\d+)?\s*(ms|s|sec|seconds|MB/s|KB/s|ops/s|/s|cmd/s|commands/sec|sandboxes/sec)\b", "N \\2", text)
    text = re.sub(r"\b\d+\.\d+\b", "F", text)
    return text


def routes(requests: list) -> list[str]:
    out = []
    for m, p in requests:
        p = re.sub(r"sbx-[0-9a-f]{12}", "{id}", p)
        p = re.sub(r"exp-[0-9a-f]{8}", "{exp}", p)
        out.append(f"{m} {p}")
    return out


def collapse(seq: list[str]) -> list[str]:
    """Polling makes run lengths timing dependent: consecutive repeats of one route count once."""
    out: list[str] = []
    for r in seq:
        if not out or out[-1] != r:
            out.append(r)
    return out


def run_one(impl: str, script: Path, timeout: int, verbose: bool) -> dict:
    sys.path.insert(0, str(REPO / "tools"))
    import local_sandbox_service as svc_mod

    httpd, svc = svc_mod.serve(0)
    port = httpd.server_address[1]
    with tempfile.TemporaryDirectory(prefix="refex_") as home:
        env = {k: v for k, v in os.environ.items() if not k.startswith("PRIME_")}
        env.update(HOME=home, PRIME_API_BASE_URL=f"http://127.0.0.1:{port}", PRIME_API_KEY="local-key", PRIME_DISABLE_VERSION_CHECK="1", PYTHONUNBUFFERED="1")
        env.pop("PYTHONPATH", None)
        boot = _BOOT.format(repo=str(REPO), ref=str(REFERENCE), sdk=SDK_DIRS)
        t0 = time.perf_counter()
        try:
            p = subprocess.run([sys.executable, "-c", boot, impl, str(script), *ARGS.get(script.stem, [])], env=env, cwd=home, capture_output=True,
                               text=True, timeout=timeout, stdin=subprocess.DEVNULL)  # fmt: skip
            code, out, err = p.returncode, p.stdout, p.stderr
        except subprocess.TimeoutExpired as e:
            code, out, err = -9, (e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or ""), "timeout"
        secs = time.perf_counter() - t0
    reqs = list(svc.requests)
    leftover = sum(1 for s in svc.sandboxes.values() if s.status != "TERMINATED")
    httpd.shutdown()
    svc.close()
    if verbose:
        print(f"--- {impl} {script.name} exit={code} {secs:.1f}s requests={len(reqs)}\n{out[-3000:]}\n{err[-1500:]}", flush=True)
    return {"exit": code, "seconds": round(secs, 2), "requests": len(reqs), "routes": routes(reqs), "stdout": out, "stderr_tail": err[-600:],
            "sandboxes_created": len(svc.sandboxes), "sandboxes_left_running": leftover}  # fmt: skip


def main(argv: list[str] | None = None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--only", nargs="*")
    ap.add_argument("--impl", choices=("ours", "reference", "both"), default="both")
    ap.add_argument("--timeout", type=int, default=600)
    ap.add_argument("--out", default=str(REPO / "profiles" / "reference_examples.json"))
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args(argv)
    ex_dir = REFERENCE / "examples"
    if not ex_dir.is_dir():
        print(json.dumps({"unavailable": f"{ex_dir} not found"}))
        return 0
    report: dict = {"examples": {}, "skipped": SKIP}
    bad = 0
    for script in sorted(ex_dir.glob("*.py")):
        if script.stem in SKIP or (a.only and script.stem not in a.only):
            continue
        arms = {impl: run_one(impl, script, a.timeout, a.verbose) for impl in (("ours", "reference") if a.impl == "both" else (a.impl,))}
        entry: dict = {impl: {k: r[k] for k in ("exit", "seconds", "requests", "sandboxes_created", "sandboxes_left_running")} for impl, r in arms.items()}
        if len(arms) == 2:
            o, r = arms["ours"], arms["reference"]
            same_out = mask(o["stdout"]) == mask(r["stdout"])
            entry.update(same_exit=o["exit"] == r["exit"], same_routes=collapse(o["routes"]) == collapse(r["routes"]),
                         same_route_multiset=sorted(set(o["routes"])) == sorted(set(r["routes"])), same_masked_stdout=same_out)  # fmt: skip
            if not same_out:
                ol, rl = mask(o["stdout"]).splitlines(), mask(r["stdout"]).splitlines()
                entry["first_stdout_difference"] = next(({"line": i, "ours": x, "reference": y} for i, (x, y) in enumerate(zip(ol, rl)) if x != y),
                                                        {"line": min(len(ol), len(rl)), "ours_lines": len(ol), "reference_lines": len(rl)})  # fmt: skip
            ok = entry["same_exit"] and o["exit"] == 0 and entry["same_route_multiset"]
        else:
            only = next(iter(arms.values()))
            ok = only["exit"] == 0
            if not ok:
                entry["stderr_tail"] = only["stderr_tail"]
        entry["ok"] = ok
        bad += not ok
        report["examples"][script.stem] = entry
        print(f"{script.stem:40s} " + json.dumps({k: v for k, v in entry.items() if k != "first_stdout_difference"}), flush=True)
    report["all_ok"] = bad == 0
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(report, indent=1, sort_keys=True) + "\n")
    print(f"{len(report['examples'])} examples, {bad} not ok")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
