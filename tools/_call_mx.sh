timeout 400 python -m pytest tests/test_kernels_gpu.py -q -k "mxfp8 or fp8 or flash_attention" 2>&1 | tail -15
timeout 200 python tools/mxfp8_bench.py > gpurun_out/mxfp8_bench.json 2> gpurun_out/mxfp8_bench.err; tail -3 gpurun_out/mxfp8_bench.err; head -c 1500 gpurun_out/mxfp8_bench.json
