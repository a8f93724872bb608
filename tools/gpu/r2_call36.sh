#!/bin/bash
# round 2, GPU call 36: cumulative open-ended windows (pair predicates, per-channel minimum / sum scans)
mkdir -p gpurun_out/r2aj
timeout 900 python -m pytest tests/test_gpu_open_window.py tests/test_gpu_leaf_predicates.py -m gpu -q > gpurun_out/r2aj/pytest.log 2>&1
grep -n "^E  \|^FAILED\|passed\|failed" gpurun_out/r2aj/pytest.log | cut -c1-400 | head -40
