#!/bin/bash
# round 2, GPU call 35: vmap rules of the scan route, smoke() with the open-ended block
mkdir -p gpurun_out/r2ai
timeout 900 python -m pytest tests/test_gpu_open_window.py tests/test_gpu_leaf_predicates.py tests/test_gpu_interface.py -m gpu -q > gpurun_out/r2ai/pytest.log 2>&1
tail -30 gpurun_out/r2ai/pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ai/smoke.log 2>&1; tail -9 gpurun_out/r2ai/smoke.log
