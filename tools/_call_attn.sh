for bkv in 64 128; do
echo "BKV=$bkv"
PB_ATTN_FWD2_BKV=$bkv timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "two_tile_kernel or rope_attention or qkv_gemm_rope" 2>&1 | tail -3 | cut -c1-300
PB_ATTN_FWD2_BKV=$bkv timeout 200 python tools/attn_bench.py 2>/dev/null | tee gpurun_out/attn_bench_v9_bkv$bkv.json | cut -c1-250
done
