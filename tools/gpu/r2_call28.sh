#!/bin/bash
# round 2, GPU call 28: open-ended formulas through the public API at 16384 x 16384 x 2 (before the host-side shortcuts)
mkdir -p gpurun_out/r2ab
timeout 600 python tools/open_bench.py > gpurun_out/r2ab/open_bench.jsonl 2> gpurun_out/r2ab/open_bench.err
cat gpurun_out/r2ab/open_bench.jsonl; tail -5 gpurun_out/r2ab/open_bench.err
