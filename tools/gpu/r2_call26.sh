#!/bin/bash
# round 2, GPU call 26: classical Until without an end beyond one tile
mkdir -p gpurun_out/r2z
timeout 900 python -m pytest tests/test_gpu_open_window.py -m gpu -q --maxfail=12 > gpurun_out/r2z/pytest_open.log 2>&1
tail -5 gpurun_out/r2z/pytest_open.log
timeout 300 python tools/scan_bench.py > gpurun_out/r2z/scan_bench.jsonl 2> gpurun_out/r2z/scan_bench.err
cat gpurun_out/r2z/scan_bench.jsonl; tail -3 gpurun_out/r2z/scan_bench.err
