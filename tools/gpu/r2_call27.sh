#!/bin/bash
# round 2, GPU call 27: chunk-boundary / degenerate-shape test of the scan kernels
mkdir -p gpurun_out/r2aa
timeout 900 python -m pytest tests/test_gpu_open_window.py -m gpu -q --maxfail=12 > gpurun_out/r2aa/pytest_open.log 2>&1
tail -5 gpurun_out/r2aa/pytest_open.log
