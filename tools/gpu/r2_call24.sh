#!/bin/bash
# round 2, GPU call 24: chunked (coalesced, one read + one write) suffix-scan kernels; their own timing
mkdir -p gpurun_out/r2x
timeout 900 python -m pytest tests/test_gpu_open_window.py -m gpu -q --maxfail=12 > gpurun_out/r2x/pytest_open.log 2>&1
tail -5 gpurun_out/r2x/pytest_open.log
timeout 300 python tools/scan_bench.py > gpurun_out/r2x/scan_bench.jsonl 2> gpurun_out/r2x/scan_bench.err
cat gpurun_out/r2x/scan_bench.jsonl; tail -3 gpurun_out/r2x/scan_bench.err
