#!/bin/bash
# Run ON the GPU box: compute-sanitizer over the kernel tests. memcheck = out-of-bounds / misaligned global+shared accesses,
# racecheck = shared-memory hazards, synccheck = invalid barrier usage. tcgen05/TMA kernels run ~50-100x slower under the tool,
# so each tool gets a focused subset; summaries land in gpurun_out/sanitizer_*.txt (copy to profiles/).
set -u
SAN=compute-sanitizer
run() { # tool, pytest -k expression, tag
  timeout 600 $SAN --tool "$1" --print-limit 5 --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -q -k "$2" -p no:cacheprovider \
    > gpurun_out/sanitizer_$3.log 2>&1
  echo "$1 [$2]: exit $? — $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_$3.log | tr '\n' ' ')" | tee -a gpurun_out/sanitizer_summary.txt
}
: > gpurun_out/sanitizer_summary.txt
run memcheck  "rmsnorm or swiglu or rope_qkv or cross_entropy" memcheck_elementwise
run memcheck  "gemm_layouts or gemm_fp32 or split_k" memcheck_gemm
run memcheck  "flash_attention or rope_attention or two_tile or dkdv_variants" memcheck_attention
run racecheck "rmsnorm or cross_entropy" racecheck_elementwise
run synccheck "gemm_pair or flash_attention or rope_attention or qkv_gemm_rope or two_tile or dkdv_variants" synccheck_tc
run memcheck  "mxfp8" memcheck_mxfp8
run synccheck "mxfp8 or swiglu_epilogue" synccheck_mxfp8
cat gpurun_out/sanitizer_summary.txt
