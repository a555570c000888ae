#!/usr/bin/env python
"""Per-kernel SASS listings of the built library → ``profiles/sass/`` (listings, not a mnemonic-count summary).

    python tools/sass_listings.py            # every kernel family, one representative instantiation each, + INDEX.md
    python tools/sass_listings.py --all      # every instantiation (≈30 MB of text: not committed)

A listing is the ``cuobjdump -sass -fun <mangled>`` output with the encoding words stripped: address + instruction per line.
INDEX.md tabulates, per listed kernel, the instruction count and the mnemonics that prove the hardware path
(UTC*MMA = tcgen05.mma, LDTM/STTM = TMEM, UTMA* = TMA, SYNCS = mbarrier, *.SYS = peer traffic, LDGMC/multimem = NVLS).
"""

from __future__ import annotations

import argparse
import re
import subprocess
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "prime_b200" / "_C" / "libprime_b200.so"
OUT = ROOT / "profiles" / "sass"
PROOF = ["UTCHMMA", "UTCQMMA", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "LDGMC", "HMMA", "MUFU.EX2"]
# representative instantiation per family (demangled-name regex → short file name); first match wins
PICK = [
    (r"gemm_bf16_kernel<0, 0, 1, 0, 0>", "gemm_bf16_pair_kmajor"),
    (r"gemm_bf16_kernel<1, 1, 1, 0, 0>", "gemm_bf16_pair_wgrad_mnmajor"),
    (r"gemm_bf16_kernel<0, 0, 1, 1, 0>", "gemm_bf16_pair_rope_epilogue"),
    (r"gemm_bf16_kernel<0, 0, 1, 2, 0>", "gemm_bf16_pair_swiglu_epilogue"),
    (r"gemm_bf16_kernel<0, 0, 1, 0, 1>", "gemm_allgather_activation_IO1"),
    (r"gemm_bf16_kernel<0, 0, 1, 0, 2>", "gemm_reduce_scatter_IO2"),
    (r"gemm_bf16_kernel<0, 0, 1, 0, 3>", "gemm_param_gather_fwd_IO3"),
    (r"gemm_bf16_kernel<0, 1, 1, 0, 3>", "gemm_param_gather_dgrad_IO3"),
    (r"gemm_bf16_kernel<0, 0, 1, 2, 3>", "gemm_param_gather_swiglu_IO3"),
    (r"gemm_bf16_kernel<0, 0, 0, 0, 0>", "gemm_bf16_single_cta"),
    (r"gemm_mxfp8_kernel<1>", "gemm_mxfp8_pair"),
    (r"quantize_mxfp8_kernel", "quantize_mxfp8"),
    (r"flash_fwd2_kernel<128>", "flash_fwd2_d128"),
    (r"flash_fwd_kernel<128>", "flash_fwd1_d128"),
    (r"bwd_dkdv_kernel<128, 1>", "flash_bwd_dkdv_d128"),
    (r"bwd_dq_kernel<128, 1>", "flash_bwd_dq_d128"),
    (r"bwd_delta_kernel", "flash_bwd_delta"),
    (r"rmsnorm_fwd_kernel<4, true>", "rmsnorm_fwd_residual"),
    (r"rmsnorm_bwd_kernel<4, true>", "rmsnorm_bwd_residual"),
    (r"swiglu_bwd_kernel", "swiglu_bwd"),
    (r"cross_entropy_kernel", "cross_entropy_fwd_bwd"),
    (r"embedding_fwd_kernel", "embedding_fwd_peer_gather"),
    (r"embedding_sort_kernel", "embedding_bwd_sort"),
    (r"embedding_bwd_kernel", "embedding_bwd_segmented"),
    (r"^grad_reduce_kernel", "grad_reduce_p2p"),
    (r"grad_reduce_segs_kernel", "grad_reduce_p2p_rows"),
    (r"adamw_push_kernel", "adamw_push"),
    (r"norm_publish_kernel", "norm_publish"),
    (r"pseudograd_quant_kernel", "outer_pseudograd_quant"),
    (r"^outer_nesterov_kernel", "outer_nesterov_int8"),
    (r"outer_nesterov_f32_kernel", "outer_nesterov_f32"),
    (r"barrier_kernel", "flag_barrier"),
    (r"mc_grad_reduce", "nvls_grad_reduce"),
    (r"mc_all_reduce", "nvls_all_reduce"),
]


def functions() -> list[tuple[str, str]]:
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    mangled = re.findall(r"Function : (\S+)", txt)
    dem = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True, check=True).stdout.splitlines()
    return [(m, d.replace("void ", "").replace("(anonymous namespace)::", "")) for m, d in zip(mangled, dem)]


def listing(mangled: str) -> list[str]:
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, str(LIB)], capture_output=True, text=True, check=True).stdout
    out = []
    for ln in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", ln)
        if m:
            out.append(f"/*{m.group(1)}*/  {m.group(2).strip()} ;")
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--all", action="store_true")
    args = ap.parse_args()
    OUT.mkdir(parents=True, exist_ok=True)
    fns = functions()
    todo: list[tuple[str, str, str]] = []
    if args.all:
        todo = [(m, d, re.sub(r"[^A-Za-z0-9_]+", "_", d.split("(")[0])[:100]) for m, d in fns]
    else:
        for pat, short in PICK:
            for m, d in fns:
                if re.search(pat, d):
                    todo.append((m, d, short))
                    break
    rows = []
    for m, d, short in todo:
        lines = listing(m)
        (OUT / f"{short}.sass").write_text(f"// {d.split('(')[0]}\n// {m}\n" + "\n".join(lines) + "\n")
        c = Counter()
        for ln in lines:
            ins = ln.split("*/", 1)[1]
            for k in PROOF:
                if re.search(rf"\b{re.escape(k)}", ins):
                    c[k] += 1
            if re.search(r"\.SYS\b", ins):
                c["*.SYS"] += 1
        rows.append((short, d.split("(")[0], len(lines), c))
    cols = PROOF + ["*.SYS"]
    idx = ["# SASS listings (`cuobjdump -sass -fun`, encodings stripped)", "",
           f"{len(fns)} kernels in `libprime_b200.so`; one representative instantiation per family listed here "
           "(`python tools/sass_listings.py --all` writes every one).", "",
           "| file | kernel | instr | " + " | ".join(cols) + " |", "|---|---|---:|" + "---:|" * len(cols)]  # fmt: skip
    for short, d, n, c in rows:
        idx.append(f"| {short}.sass | `{d}` | {n} | " + " | ".join(str(c.get(k, "")) for k in cols) + " |")
    (OUT / "INDEX.md").write_text("\n".join(idx) + "\n")
    print(f"wrote {len(rows)} listings to {OUT}")


if __name__ == "__main__":
    main()
