timeout 600 python -m pytest tests/test_multigpu.py -q -x -k "fused_allgather or sequence_parallel" 2>&1 | tail -15
