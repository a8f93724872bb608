#!/bin/bash
# round 2, GPU call 23: open-ended Always / Eventually beyond one tile (suffix-scan kernels, csrc/scan.cu)
mkdir -p gpurun_out/r2w
timeout 900 python -m pytest tests/test_gpu_open_window.py -m gpu -q --maxfail=12 > gpurun_out/r2w/pytest_open.log 2>&1
tail -40 gpurun_out/r2w/pytest_open.log
