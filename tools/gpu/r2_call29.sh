#!/bin/bash
# round 2, GPU call 29: root-leaf shortcut + channel remap on the host side: full GPU suite, open-ended formulas through the public API
mkdir -p gpurun_out/r2ac
timeout 1500 python -m pytest tests -m gpu -q -n 4 > gpurun_out/r2ac/pytest.log 2>&1
tail -6 gpurun_out/r2ac/pytest.log
timeout 600 python tools/open_bench.py > gpurun_out/r2ac/open_bench.jsonl 2> gpurun_out/r2ac/open_bench.err
cut -c1-420 gpurun_out/r2ac/open_bench.jsonl; tail -5 gpurun_out/r2ac/open_bench.err
