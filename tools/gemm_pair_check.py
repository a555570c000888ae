"""Validate and time the CTA-pair (cta_group::2) GEMM scheduler against the single-CTA one and cuBLAS.

    python tools/gemm_pair_check.py            # numerics on awkward shapes, then the Llama-1B shapes timed both ways
"""

import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from prime_b200 import ops  # noqa: E402


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


def timeit(fn, flush, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    bad = []
    ops.set_gemm_pair_mode(True)
    for M, N, K in [(256, 256, 64), (256, 512, 256), (384, 1000, 136), (2048, 2048, 2048), (136, 264, 72), (4096, 4096, 1024), (640, 256, 4096), (16384, 6144, 2048)]:
        for a_mn in (False, True):
            for b_mn in (False, True):
                A = torch.randn(M, K, device=dev, dtype=torch.bfloat16) * 0.1
                B = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.1
                ref = A.float() @ B.float().t()
                a = A.t().contiguous() if a_mn else A
                b = B.t().contiguous() if b_mn else B
                out = ops.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn)
                torch.cuda.synchronize()
                e = rel(out, ref)
                if not e < 1e-2:
                    bad.append((M, N, K, a_mn, b_mn, e))
    # fp32 accumulate epilogue (TMA reduce-add)
    M, N, K = 512, 768, 1024
    A = torch.randn(K, M, device=dev, dtype=torch.bfloat16)
    B = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
    C = torch.randn(M, N, device=dev, dtype=torch.float32)
    ref = C + A.float().t() @ B.float()
    ops.gemm(A, B, a_mn_major=True, b_mn_major=True, out=C, accumulate=True)
    torch.cuda.synchronize()
    if not rel(C, ref) < 2e-3:
        bad.append(("fp32acc", rel(C, ref)))
    print(json.dumps({"pair_numerics_failures": bad}))
    if bad:
        sys.exit(1)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    T = 16384
    shapes = [("qkv fwd", T, 6144, 2048, False, False), ("wo fwd", T, 2048, 2048, False, False), ("w13 fwd", T, 11264, 2048, False, False),
              ("w2 fwd", T, 2048, 5632, False, False), ("logits fwd", T, 32000, 2048, False, False), ("w13 dgrad", T, 2048, 11264, False, True),
              ("w13 wgrad", 11264, 2048, T, True, True), ("w2 wgrad", 2048, 5632, T, True, True), ("square 8192", 8192, 8192, 8192, False, False)]  # fmt: skip
    rows = []
    for name, M, N, K, amn, bmn in shapes:
        A = torch.randn((K, M) if amn else (M, K), device=dev, dtype=torch.bfloat16)
        B = torch.randn((K, N) if bmn else (N, K), device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        res = {}
        for mode in (False, True):
            ops.set_gemm_pair_mode(mode)
            res["pair" if mode else "single"] = timeit(lambda: ops.gemm(A, B, a_mn_major=amn, b_mn_major=bmn, out=out), flush)
        At = A.t() if amn else A
        Bt = B if bmn else B.t()
        res["cublas"] = timeit(lambda: torch.matmul(At, Bt, out=out), flush)
        fl = 2.0 * M * N * K
        rows.append({"shape": name, "M": M, "N": N, "K": K, **{f"{k}_ms": round(v, 4) for k, v in res.items()},
                     **{f"{k}_tflops": round(fl / v / 1e9, 1) for k, v in res.items()}})  # fmt: skip
        print(json.dumps(rows[-1]), flush=True)
    ops.set_gemm_pair_mode(None)
    print(json.dumps({"rows": rows}))


if __name__ == "__main__":
    main()
