"""MXFP8 GEMM (tcgen05 kind::mxf8f6f4.block_scale) vs the bf16 tcgen05 GEMM on the Llama-1B projection shapes.

Device-timed with CUDA events, median of 20, L2 flushed between iterations. `python tools/mxfp8_bench.py > out.json`"""

import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from prime_b200 import ops  # noqa: E402


def timeit(fn, flush, iters=20, warm=3):
    for _ in range(warm):
        fn()
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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    T = 16384
    shapes = [("qkv", T, 6144, 2048), ("wo", T, 2048, 2048), ("w13", T, 11264, 2048), ("w2", T, 2048, 5632), ("dgrad_w13", T, 2048, 11264)]
    rows = []
    for name, M, N, K in shapes:
        a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        b = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        aq, asf = ops.quantize_mxfp8(a)
        bq, bsf = ops.quantize_mxfp8(b)
        ops.set_mxfp8_pair_mode(False)
        t_fp8_single = timeit(lambda: ops.gemm_mxfp8(aq, asf, bq, bsf, out=out), flush)
        ops.set_mxfp8_pair_mode(True)
        t_fp8 = timeit(lambda: ops.gemm_mxfp8(aq, asf, bq, bsf, out=out), flush)
        t_bf16 = timeit(lambda: ops.gemm(a, b, out=out), flush)
        t_qa = timeit(lambda: ops.quantize_mxfp8(a), flush)
        t_qb = timeit(lambda: ops.quantize_mxfp8(b), flush)
        t_qbt = timeit(lambda: ops.quantize_mxfp8(b, transpose=True), flush) if N % 128 == 0 else None
        fl = 2.0 * M * N * K
        rows.append({"shape": name, "M": M, "N": N, "K": K, "mxfp8_ms": round(t_fp8, 4), "mxfp8_tflops": round(fl / t_fp8 / 1e9, 1),
                     "mxfp8_single_cta_ms": round(t_fp8_single, 4), "mxfp8_single_cta_tflops": round(fl / t_fp8_single / 1e9, 1),
                     "bf16_ms": round(t_bf16, 4), "bf16_tflops": round(fl / t_bf16 / 1e9, 1), "speedup": round(t_bf16 / t_fp8, 3),
                     "quantize_act_ms": round(t_qa, 4), "quantize_act_GBps": round(M * K * 3 / t_qa / 1e6, 1),
                     "quantize_w_ms": round(t_qb, 4), "quantize_w_transposed_ms": None if t_qbt is None else round(t_qbt, 4)})  # fmt: skip
    print(json.dumps({"rows": rows, "note": "fp8 dense peak 4.5 PFLOP/s nominal; bf16 2.25"}, indent=1))


if __name__ == "__main__":
    main()
