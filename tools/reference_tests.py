#!/usr/bin/env python
"""Run the reference's OWN test files, unmodified, against this package — and against the reference itself as the control arm.

    python tools/reference_tests.py                      # both arms, all five packages → profiles/reference_tests.json
    python tools/reference_tests.py --impl ours --packages prime-tunnel -v

"ours": ``prime_b200.compat.install()`` registers ``prime_cli`` / ``prime_sandboxes`` / ``prime_evals`` / ``prime_tunnel`` / ``prime_mcp``
as aliases of ``prime_b200.platform.*`` and pytest then collects ``/root/reference/packages/<pkg>/tests`` as they are.  "reference": the
same files with the reference's ``src`` directories first on ``sys.path``.  Each (arm, package) runs in its own process with a scratch
``HOME`` and no ``PRIME_*`` variables, so neither arm sees credentials and neither can touch ``~/.prime``.

Tests that need the network (live sandbox API) fail the same way in both arms; what matters is the DIFFERENCE: a test the control arm
passes and ours does not.  The summary lists exactly those, with the first line of the failure — most are white-box tests that
monkey-patch private module attributes of the reference's file layout (``prime_cli.commands.evals._fetch_logs`` …).

``pytest-asyncio`` is not installed in this image; a ten-line shim below runs ``async def`` tests with ``asyncio.run`` in both arms.
"""

from __future__ import annotations

import argparse
import asyncio
import inspect
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
REFERENCE = Path(os.environ.get("PRIME_REFERENCE_ROOT", "/root/reference"))
PACKAGES = ("prime", "prime-sandboxes", "prime-evals", "prime-tunnel", "prime-mcp-server")


# ---------------------------------------------------------------------------------------------------------------- inner: one (arm, package)
class _Recorder:
    def __init__(self) -> None:
        self.outcomes: dict[str, str] = {}
        self.why: dict[str, str] = {}

    def pytest_runtest_logreport(self, report) -> None:
        bad = report.failed
        if report.when == "call" or bad or (report.when == "setup" and report.skipped):
            prev = self.outcomes.get(report.nodeid)
            outcome = "error" if (bad and report.when != "call") else report.outcome
            if prev in (None, "passed"):
                self.outcomes[report.nodeid] = outcome
            if bad:
                self.why[report.nodeid] = _last_line(str(report.longrepr))

    def pytest_collectreport(self, report) -> None:
        if report.failed:
            self.outcomes[report.nodeid] = "collect_error"
            self.why[report.nodeid] = _last_line(str(report.longrepr))


class _AsyncShim:
    """Stand-in for pytest-asyncio (absent here): a coroutine test function is driven to completion with ``asyncio.run``."""

    @staticmethod
    def pytest_configure(config) -> None:
        config.addinivalue_line("markers", "asyncio: run the coroutine test with asyncio.run")

    @staticmethod
    def pytest_pyfunc_call(pyfuncitem):
        if inspect.iscoroutinefunction(pyfuncitem.obj):
            kwargs = {a: pyfuncitem.funcargs[a] for a in pyfuncitem._fixtureinfo.argnames}
            asyncio.run(pyfuncitem.obj(**kwargs))
            return True
        return None


def _last_line(text: str) -> str:
    lines = [ln.strip() for ln in text.strip().splitlines() if ln.strip()]
    picked = next((ln for ln in reversed(lines) if ln.startswith("E ")), lines[-1] if lines else "")
    return picked.lstrip("E ").strip()[:300]


def _inner(impl: str, package: str, verbose: bool) -> int:
    import pytest

    if impl == "ours":
        sys.path.insert(0, str(REPO))
        from prime_b200 import compat

        compat.install()
    else:
        for p in PACKAGES:
            sys.path.insert(0, str(REFERENCE / "packages" / p / "src"))
    rec = _Recorder()
    tests = REFERENCE / "packages" / package / "tests"
    argv = [str(tests), "-q", "-p", "no:cacheprovider", "--rootdir", os.environ["HOME"], "-o", "addopts=", "--tb=short" if verbose else "--tb=no",
            "--continue-on-collection-errors", "--no-header", "-W", "ignore"]  # fmt: skip
    code = pytest.main(argv, plugins=[rec, _AsyncShim()])
    Path(os.environ["_REFTEST_OUT"]).write_text(json.dumps({"exit": int(code), "outcomes": rec.outcomes, "why": rec.why}))
    return 0


# ---------------------------------------------------------------------------------------------------------------- outer: orchestrate + compare
def run_arm(impl: str, package: str, verbose: bool = False, timeout: int = 1800) -> dict:
    with tempfile.TemporaryDirectory(prefix="reftest_") as home:
        out = Path(home) / "result.json"
        env = {k: v for k, v in os.environ.items() if not k.startswith("PRIME_") or k == "PRIME_REFERENCE_ROOT"}
        env.update(HOME=home, _REFTEST_OUT=str(out), PRIME_DISABLE_VERSION_CHECK="1", PYTHONDONTWRITEBYTECODE="1", COLUMNS="200")
        env.pop("PYTHONPATH", None)
        proc = subprocess.run([sys.executable, __file__, "--inner", impl, package] + (["-v"] if verbose else []), env=env, cwd=home,
                              capture_output=not verbose, text=True, timeout=timeout)  # fmt: skip
        if not out.exists():
            return {"exit": proc.returncode, "outcomes": {}, "why": {}, "crashed": (proc.stderr or "")[-500:]}
        return json.loads(out.read_text())


def counts(outcomes: dict[str, str]) -> dict[str, int]:
    c: dict[str, int] = {}
    for o in outcomes.values():
        c[o] = c.get(o, 0) + 1
    return dict(sorted(c.items()))


def compare(ours: dict, ref: dict) -> dict:
    """Tests the control arm passes that ours does not — the only ones that say something about this package."""
    gaps = {}
    for nodeid, o in ref["outcomes"].items():
        if o == "passed" and ours["outcomes"].get(nodeid) != "passed":
            mine = ours["outcomes"].get(nodeid)
            if mine is None:  # the whole file failed to collect in our arm
                fname = nodeid.split("::")[0]
                mine, why = "collect_error", ours["why"].get(fname, "not collected")
            else:
                why = ours["why"].get(nodeid, "")
            gaps[nodeid] = {"ours": mine, "why": why}
    only_ours = sorted(n for n, o in ours["outcomes"].items() if o == "passed" and ref["outcomes"].get(n) not in ("passed", None))
    return {"reference_passes_ours_does_not": gaps, "ours_passes_reference_does_not": only_ours}


def main(argv: list[str] | None = None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--inner", nargs=2, metavar=("IMPL", "PACKAGE"), help=argparse.SUPPRESS)
    ap.add_argument("--impl", choices=("ours", "reference", "both"), default="both")
    ap.add_argument("--packages", nargs="*", default=list(PACKAGES))
    ap.add_argument("--out", default=str(REPO / "profiles" / "reference_tests.json"))
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args(argv)
    if a.inner:
        return _inner(a.inner[0], a.inner[1], a.verbose)
    if not (REFERENCE / "packages").is_dir():
        print(json.dumps({"unavailable": f"{REFERENCE}/packages not found"}))
        return 0
    report: dict = {"reference_root": str(REFERENCE), "packages": {}}
    for pkg in a.packages:
        entry: dict = {}
        arms = {}
        for impl in ("ours", "reference") if a.impl == "both" else (a.impl,):
            arms[impl] = run_arm(impl, pkg, a.verbose)
            entry[impl] = {"counts": counts(arms[impl]["outcomes"]), "pytest_exit": arms[impl]["exit"]}
            if arms[impl].get("crashed"):
                entry[impl]["crashed"] = arms[impl]["crashed"]
        if len(arms) == 2:
            entry.update(compare(arms["ours"], arms["reference"]))
        elif "ours" in arms:
            entry["not_passed"] = {n: {"ours": o, "why": arms["ours"]["why"].get(n, "")} for n, o in arms["ours"]["outcomes"].items()
                                   if o not in ("passed", "skipped")}  # fmt: skip
        report["packages"][pkg] = entry
        line = {k: entry[k]["counts"] for k in arms}
        print(f"{pkg:18s} {json.dumps(line)}" + (f"  gaps={len(entry['reference_passes_ours_does_not'])}" if len(arms) == 2 else ""), flush=True)
    if a.impl == "both":
        tot = {arm: sum(report["packages"][p][arm]["counts"].get("passed", 0) for p in report["packages"]) for arm in ("ours", "reference")}
        report["total_passed"] = tot
        report["total_gaps"] = sum(len(report["packages"][p]["reference_passes_ours_does_not"]) for p in report["packages"])
        print(f"passed: ours {tot['ours']} / reference {tot['reference']}; reference-only passes: {report['total_gaps']}")
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(report, indent=1, sort_keys=True) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
