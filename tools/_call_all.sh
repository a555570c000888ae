mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.txt
python bench.py --steps 8 --warmup 3 2>/dev/null | tee gpurun_out/bench_final2.json | cut -c1-200
bash tools/sanitize.sh 2>&1 | tail -9
