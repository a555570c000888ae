"""Training metrics: throughput/MFU meter, JSONL sink, optional Prometheus endpoint.

MFU is reported against the MEASURED sustained cuBLAS bf16 throughput in ``MEASURED_PEAKS.json`` (driver-written) —
the same denominator the profiling recipe prescribes — falling back to the nominal 2.25 PFLOP/s dense bf16 peak.
"""

from __future__ import annotations

import json
import time
from collections import deque
from pathlib import Path
from typing import Any

NOMINAL_BF16_TFLOPS = 2250.0


def measured_peaks(root: Path | None = None) -> dict[str, Any]:
    p = (root or Path(__file__).resolve().parents[2]) / "MEASURED_PEAKS.json"
    try:
        return json.loads(p.read_text())
    except (OSError, json.JSONDecodeError):
        return {}


def peak_tflops(kind: str = "sustained") -> float:
    m = measured_peaks()
    if kind == "sustained" and m.get("bf16_tflops_sustained"):
        return float(m["bf16_tflops_sustained"])
    return float(m.get("bf16_tflops") or NOMINAL_BF16_TFLOPS)


class Throughput:
    """Sliding-window tokens/s and MFU. ``update`` takes the tokens of one step and the wall time it took."""

    def __init__(self, flops_per_token: float, n_gpus: int, window: int = 20):
        self.fpt, self.n_gpus = flops_per_token, max(1, n_gpus)
        self.win: deque[tuple[int, float]] = deque(maxlen=window)
        self.total_tokens = 0
        self.peak = peak_tflops() * 1e12

    def update(self, tokens: int, seconds: float) -> None:
        self.win.append((tokens, seconds))
        self.total_tokens += tokens

    @property
    def tokens_per_s(self) -> float:
        t = sum(s for _, s in self.win)
        return sum(n for n, _ in self.win) / t if t > 0 else 0.0

    @property
    def mfu(self) -> float:
        return self.tokens_per_s * self.fpt / (self.peak * self.n_gpus)


class JsonlSink:
    def __init__(self, path: str | Path | None):
        self.f = None
        if path:
            Path(path).parent.mkdir(parents=True, exist_ok=True)
            self.f = open(path, "a", buffering=1)

    def write(self, row: dict[str, Any]) -> None:
        if self.f is not None:
            self.f.write(json.dumps(row) + "\n")

    def close(self) -> None:
        if self.f is not None:
            self.f.close()
            self.f = None


class PrometheusSink:
    """Gauges for the last logged step (prometheus_client is in the image; silently disabled if the port is taken)."""

    def __init__(self, port: int | None):
        self.gauges: dict[str, Any] = {}
        self.ok = False
        if port:
            try:
                import prometheus_client as pc

                pc.start_http_server(port)
                self.pc, self.ok = pc, True
            except Exception:
                self.ok = False

    def write(self, row: dict[str, Any]) -> None:
        if not self.ok:
            return
        for k, v in row.items():
            if isinstance(v, (int, float)) and not isinstance(v, bool):
                g = self.gauges.get(k)
                if g is None:
                    g = self.gauges[k] = self.pc.Gauge(f"prime_b200_{k}", k)
                g.set(v)


class StepTimer:
    def __init__(self) -> None:
        self.t = time.perf_counter()

    def lap(self) -> float:
        now = time.perf_counter()
        dt, self.t = now - self.t, now
        return dt

    def exclude(self, seconds: float) -> None:
        """Take ``seconds`` of non-training work (a validation pass) out of the lap in progress."""
        self.t += max(0.0, seconds)
