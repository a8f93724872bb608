#!/bin/bash
# round 2, GPU call 34: the GPU suite exactly as the driver runs it (serial, -x), with the slowest tests listed; smoke; default bench; scan bench
mkdir -p gpurun_out/r2ah
( time timeout 2400 python -m pytest tests/ -x -q -m gpu --durations=25 > gpurun_out/r2ah/pytest_serial.log 2>&1 ) 2> gpurun_out/r2ah/pytest_serial.time
tail -40 gpurun_out/r2ah/pytest_serial.log; cat gpurun_out/r2ah/pytest_serial.time
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ah/smoke.log 2>&1; tail -6 gpurun_out/r2ah/smoke.log
timeout 600 python bench.py > gpurun_out/r2ah/bench_default.json 2> gpurun_out/r2ah/bench_default.err; head -c 400 gpurun_out/r2ah/bench_default.json; echo
timeout 300 python tools/scan_bench.py > gpurun_out/r2ah/scan_bench.jsonl 2> gpurun_out/r2ah/scan_bench.err; cut -c1-200 gpurun_out/r2ah/scan_bench.jsonl
