timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "flash_attention or two_tile or dkdv_variants or rope_attention" 2>&1 | tail -2
timeout 600 compute-sanitizer --tool synccheck --print-limit 5 --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -q -k "gemm_pair or flash_attention or rope_attention or qkv_gemm_rope or two_tile or dkdv_variants" -p no:cacheprovider > gpurun_out/sanitizer_synccheck_tc.log 2>&1
echo "synccheck [tensor-core kernels]: exit $? — $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_synccheck_tc.log | tr '\n' ' ')" | tee gpurun_out/sanitizer_synccheck_tc_summary.txt
timeout 200 python tools/attn_bench.py 2>/dev/null | tee gpurun_out/attn_bench_final.json | cut -c1-200
