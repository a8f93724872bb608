#!/bin/bash
# round 2, GPU call 37: random formulas with open-ended operators, scan route against the in-kernel route (and the oracle for cumulative)
mkdir -p gpurun_out/r2ak
timeout 1200 python -m pytest tests/test_gpu_open_window.py -m gpu -q -k "random_formulas" --durations=5 > gpurun_out/r2ak/pytest.log 2>&1
grep -n "^E  \|^FAILED\|passed\|failed\|s call" gpurun_out/r2ak/pytest.log | cut -c1-600 | head -50
