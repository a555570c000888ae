"""Pipeline trace of the dK/dV backward kernel: CTA (0,0) stamps clock64() at loader / MMA / math-group events; this prints
per-tile phase durations (cycles) so the critical path is read off directly instead of guessed.

    python tools/attn_trace.py [--S 1024] > gpurun_out/attn_trace.json
"""

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from prime_b200.ops import _lib  # noqa: E402
from prime_b200.ops import attention_native as A  # noqa: E402

ROLES = ["loader", "mma", "math0", "math1"]
CODES = {"loader": {1: "issue_load"}, "mma": {1: "q_landed", 2: "issue_S_dP", 3: "issue_dV_dK"},
         "math0": {1: "wait_S", 2: "S_ready", 3: "tmem_loaded", 4: "p_stage_free", 5: "done", 6: "math_done"},
         "math1": {1: "wait_S", 2: "S_ready", 3: "tmem_loaded", 4: "p_stage_free", 5: "done", 6: "math_done"}}  # fmt: skip


def summarize_dq(ev):
    """dQ kernel, CTA (0,0) = the last query block (sees every kv tile): roles loader / MMA / math group 0 / 1."""
    names = {0: {1: "issue_load"}, 1: {1: "kv_landed", 2: "issue_S_dP", 3: "issue_dQ"},
             2: {1: "wait_S", 2: "S_ready", 3: "tmem_loaded", 4: "ds_buffer_free", 5: "done"},
             3: {1: "wait_S", 2: "S_ready", 3: "tmem_loaded", 4: "ds_buffer_free", 5: "done"}}
    firsts = [int(ev[r, 0, 2]) for r in range(4) if ev[r, 0, 0] != 0]
    if not firsts:
        return None
    t0 = min(firsts)
    by = {}
    last = 0
    for r in range(4):
        for k in range(256):
            code, tile, clk = (int(x) for x in ev[r, k])
            if code == 0:
                continue
            role = ["loader", "mma", "math0", "math1"][r]
            by.setdefault(tile, {})[f"{role}.{names[r].get(code, code)}"] = clk - t0
            last = max(last, clk - t0)
    rows = []
    for tile in sorted(by):
        d = by[tile]
        g = "math0" if tile % 2 == 0 else "math1"
        rows.append({"tile": tile, "load_issue": d.get("loader.issue_load"), "kv_landed": d.get("mma.kv_landed"), "S_issue": d.get("mma.issue_S_dP"),
                     "S_ready": d.get(f"{g}.S_ready"), "wait_for_S": d.get(f"{g}.S_ready", 0) - d.get(f"{g}.wait_S", 0),
                     "tmem_load": d.get(f"{g}.tmem_loaded", 0) - d.get(f"{g}.S_ready", 0),
                     "wait_ds_buffer": d.get(f"{g}.ds_buffer_free", 0) - d.get(f"{g}.tmem_loaded", 0),
                     "math_store": d.get(f"{g}.done", 0) - d.get(f"{g}.ds_buffer_free", 0),
                     "S_issue_to_ready": d.get(f"{g}.S_ready", 0) - d.get("mma.issue_S_dP", 0),
                     "done_to_dQ_issue": d.get("mma.issue_dQ", 0) - d.get(f"{g}.done", 0)})  # fmt: skip
    timeline = []
    for r in range(4):
        for k in range(256):
            code, tile, clk = (int(x) for x in ev[r, k])
            if code == 0:
                continue
            timeline.append((clk - t0, ["loader", "mma", "math0", "math1"][r], str(names[r].get(code, code)), tile))
    timeline.sort()
    return {"total_cycles": last, "tiles": len(rows), "cycles_per_tile": round(last / max(1, len(rows))), "per_tile": rows,
            "timeline": [{"cyc": c, "role": ro, "event": e, "tile": t} for c, ro, e, t in timeline]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--S", type=int, default=1024)
    ap.add_argument("--B", type=int, default=16)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, S, H, D = a.B, a.S, 16, 128
    qkv = (torch.randn(B, S, 3 * H * D, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
    out = A.flash_attention_qkv(qkv, H, H, True)
    dout = torch.randn_like(out)
    out.backward(dout, retain_graph=True)  # warm-up
    buf = torch.zeros(8 * 256 * 3, dtype=torch.int64, device=dev)
    lib = _lib.load()
    lib.pb_flash_attn_bwd_set_trace(buf.data_ptr())
    qkv.grad = None
    out.backward(dout)
    torch.cuda.synchronize()
    lib.pb_flash_attn_bwd_set_trace(None)
    ev_all = buf.view(8, 256, 3).cpu().numpy()
    dq = summarize_dq(ev_all[4:])
    ev = ev_all[:4]
    t0 = min(int(ev[r, 0, 2]) for r in range(4) if ev[r, 0, 0] != 0)
    timeline = []
    for r, role in enumerate(ROLES):
        for k in range(256):
            code, tile, clk = (int(x) for x in ev[r, k])
            if code == 0:
                continue
            timeline.append({"role": role, "event": CODES[role].get(code, str(code)), "tile": tile, "cyc": clk - t0})
    timeline.sort(key=lambda e: e["cyc"])
    # per-tile phase summary
    by = {}
    for e in timeline:
        by.setdefault(e["tile"], {})[f"{e['role']}.{e['event']}"] = e["cyc"]
    rows = []
    for tile in sorted(by):
        d = by[tile]
        g = "math0" if tile % 2 == 0 else "math1"
        rows.append({"tile": tile, "S_issue": d.get("mma.issue_S_dP"), "S_ready": d.get(f"{g}.S_ready"),
                     "wait_for_S": (d.get(f"{g}.S_ready", 0) - d.get(f"{g}.wait_S", 0)),
                     "tmem_load": (d.get(f"{g}.tmem_loaded", 0) - d.get(f"{g}.S_ready", 0)),
                     "math": (d.get(f"{g}.math_done", 0) - d.get(f"{g}.tmem_loaded", 0)),
                     "wait_p_buffer": (d.get(f"{g}.p_stage_free", 0) - d.get(f"{g}.math_done", 0)),
                     "math_store": (d.get(f"{g}.done", 0) - d.get(f"{g}.p_stage_free", 0)),
                     "S_issue_to_ready": (d.get(f"{g}.S_ready", 0) - d.get("mma.issue_S_dP", 0)),
                     "done_to_dVdK_issue": (d.get("mma.issue_dV_dK", 0) - d.get(f"{g}.done", 0)),
                     "load_issue": d.get("loader.issue_load"), "q_landed": d.get("mma.q_landed")})  # fmt: skip
    span = timeline[-1]["cyc"] if timeline else 0
    print(json.dumps({"shape": {"B": B, "S": S, "H": H, "D": D}, "cta": "kv block 0 (sees every query tile: the longest CTA)", "total_cycles": span,
                      "tiles": len(rows), "cycles_per_tile": round(span / max(1, len(rows))), "per_tile": rows, "dq_kernel": dq}, indent=1))  # fmt: skip


if __name__ == "__main__":
    main()
