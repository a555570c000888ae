#!/usr/bin/env python
"""BASELINE config 5 on real GPUs: W elastic DiLoCo workers × G GPUs, kill -9 one worker mid-run, let the supervisor respawn it.

    python tools/elastic_gpu_demo.py --workers 4 --gpus 2 --model 1B --H 10 --steps 60 --kill-at 22 --out gpurun_out/elastic_4_3_4.json

Drives the public run manager (``python -m prime_b200.launch run --workers W --gpus G --elastic --respawn 2 …``), watches the
per-worker metrics JSONL, SIGKILLs the process group of the victim once every worker has passed ``--kill-at``, waits for the
run to finish and writes a summary: membership and job tokens/s before / during / after the drop, the outer-step device time
and bytes on the wire, whether the fused NVLink exchange was used, and the parameter hashes of all workers at the last outer
step (they must be identical — the rejoined worker received the live checkpoint from a survivor).
"""

from __future__ import annotations

import argparse
import json
import os
import signal
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def read_jsonl(p: Path) -> list[dict]:
    try:
        return [json.loads(x) for x in p.read_text().splitlines() if x.strip()]
    except FileNotFoundError:
        return []


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--model", default="1B")
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--micro-bs", type=int, default=16)
    ap.add_argument("--batch", type=int, default=0, help="sequences per worker per step (default micro_bs × gpus)")
    ap.add_argument("--H", type=int, default=10)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--kill-at", type=int, default=22)
    ap.add_argument("--victim", default="w2")
    ap.add_argument("--timeout", type=float, default=900)
    ap.add_argument("--out", default="gpurun_out/elastic_demo.json")
    a = ap.parse_args()

    runs = Path(os.environ.get("PRIME_B200_RUNS_DIR", str(ROOT / "gpurun_out" / "runs")))
    env = {**os.environ, "PRIME_B200_RUNS_DIR": str(runs), "PYTHONPATH": str(ROOT)}
    batch = a.batch or a.micro_bs * a.gpus
    targs = ["--name_model", a.model, "--data.seq_length", str(a.seq), "--optim.batch_size", str(batch), "--train.micro_bs", str(a.micro_bs),
             "--optim.warmup_steps", "5", "--optim.total_steps", str(a.steps), "--diloco.inner_steps", str(a.H), "--mesh.num_workers", str(a.workers),
             "--mesh.heartbeat_interval_s", "0.5", "--mesh.heartbeat_timeout_s", "8", "--train.log_model_hash", "true", "--train.attn_impl", "native"]  # fmt: skip
    cmd = [sys.executable, "-m", "prime_b200.launch", "run", "--workers", str(a.workers), "--gpus", str(a.gpus), "--elastic", "--respawn", "2",
           "--detach", "--", *targs]  # fmt: skip
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    if out.returncode != 0:
        raise SystemExit(out.stdout + out.stderr)
    info = json.loads(out.stdout.strip().splitlines()[-1])
    rdir = Path(info["dir"])
    names = [f"w{i}" for i in range(a.workers)]
    t0 = time.time()
    killed_at = None
    victim_pid = None
    events = []
    while time.time() - t0 < a.timeout:
        st = json.loads((rdir / "status.json").read_text()) if (rdir / "status.json").exists() else {}
        steps = {n: (read_jsonl(rdir / f"metrics-{n}.jsonl") or [{}])[-1].get("step", 0) for n in names}
        if killed_at is None and steps and min(steps.values()) >= a.kill_at:
            for w in st.get("workers", []):
                if w["name"] == a.victim and w.get("pid"):
                    victim_pid = w["pid"]
                    from prime_b200.launch import signal_tree

                    signal_tree(victim_pid, signal.SIGKILL)  # torchrun AND its ranks (they live in their own sessions)
                    killed_at = time.time()
                    events.append({"t": round(killed_at - t0, 2), "event": f"SIGKILL process group of {a.victim} (pid {victim_pid}) at steps {steps}"})
        if st.get("state") in ("COMPLETED", "FAILED", "STOPPED"):
            events.append({"t": round(time.time() - t0, 2), "event": f"run {st['state']}"})
            break
        time.sleep(0.5)
    else:
        subprocess.run([sys.executable, "-m", "prime_b200.launch", "stop", info["run"], "--force"], cwd=ROOT, env=env)
        events.append({"t": round(time.time() - t0, 2), "event": "timeout: stopped"})

    rows = {n: read_jsonl(rdir / f"metrics-{n}.jsonl") for n in names}
    st = json.loads((rdir / "status.json").read_text())
    timeline = []
    for n in names:
        for r in rows[n]:
            timeline.append({"worker": n, "step": r["step"], "time": r["time"], "workers": r["workers"], "tokens_per_s": r["tokens_per_s"],
                             "outer": r.get("outer"), "outer_s": r.get("outer_s"), "outer_bytes": r.get("outer_bytes"), "param_hash": r.get("param_hash"),
                             "loss": r["loss"]})  # fmt: skip
    # job tokens/s as seen by a survivor (w0): mean of the records in each membership phase
    w0 = rows["w0"]
    phases: dict[str, list[float]] = {}
    seen_drop = False
    for r in w0:
        if r["workers"] < a.workers:
            seen_drop = True
        key = f"{r['workers']}_workers_" + ("before" if not seen_drop else ("during" if r["workers"] < a.workers else "after"))
        phases.setdefault(key, []).append(r["tokens_per_s"])
    last_outer = {n: [r for r in rows[n] if r.get("outer")][-1] if any(r.get("outer") for r in rows[n]) else None for n in names}
    hashes = {n: (lo or {}).get("param_hash") for n, lo in last_outer.items()}
    logs = {n: (rdir / "logs" / f"{n}.log").read_text(errors="replace") if (rdir / "logs" / f"{n}.log").exists() else "" for n in names}
    summary = {
        "config": {"workers": a.workers, "gpus_per_worker": a.gpus, "model": a.model, "seq": a.seq, "micro_bs": a.micro_bs, "H": a.H, "steps": a.steps,
                   "victim": a.victim, "kill_at_step": a.kill_at},
        "state": st.get("state"),
        "workers": st.get("workers"),
        "events": events,
        "fused_nvlink_exchange": {n: "fused NVLink outer exchange" in logs[n] for n in names},
        "elastic_log": {n: [ln.split("INFO ", 1)[-1] for ln in logs[n].splitlines() if "elastic:" in ln][-12:] for n in names},
        "job_tokens_per_s_by_phase (w0's view, mean)": {k: round(sum(v) / len(v), 1) for k, v in phases.items()},
        "membership_seen_by_w0": [(r["step"], r["workers"]) for r in w0 if r.get("outer")],
        "outer_steps_w0": [{"step": r["step"], "workers": r["workers"], "outer_ms": round(1e3 * r["outer_s"], 2), "MiB_on_wire": round(r["outer_bytes"] / 2**20, 1)}
                           for r in w0 if r.get("outer")],
        "final_step": {n: (rows[n][-1]["step"] if rows[n] else None) for n in names},
        "param_hash_at_last_outer_step": hashes,
        "hashes_identical": len({h for h in hashes.values() if h is not None}) == 1 and all(h is not None for h in hashes.values()),
        "timeline": timeline,
    }  # fmt: skip
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(summary, indent=1))
    print(json.dumps({k: v for k, v in summary.items() if k != "timeline"}, indent=1))
    for n in names:
        tail = Path(a.out).with_name(Path(a.out).stem + f"_{n}.log")
        tail.write_text(logs[n][-6000:])


if __name__ == "__main__":
    main()
