#!/bin/bash
# round 2, GPU call 25: full GPU suite after the ABI-2 change, scan kernels timed alone + their DRAM traffic, smoke, default bench
mkdir -p gpurun_out/r2y
timeout 1500 python -m pytest tests -m gpu -q -n 4 > gpurun_out/r2y/pytest.log 2>&1
tail -4 gpurun_out/r2y/pytest.log
timeout 300 python tools/scan_bench.py > gpurun_out/r2y/scan_bench.jsonl 2> gpurun_out/r2y/scan_bench.err
cat gpurun_out/r2y/scan_bench.jsonl; tail -3 gpurun_out/r2y/scan_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2y/scan_ncu.csv python tools/scan_bench.py --iters 1 --warm 0 > gpurun_out/r2y/scan_ncu.log 2>&1
grep -c open_window gpurun_out/r2y/scan_ncu.csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y/smoke.log 2>&1; tail -6 gpurun_out/r2y/smoke.log
timeout 600 python bench.py > gpurun_out/r2y/bench_default.json 2> gpurun_out/r2y/bench_default.err; tail -c 1500 gpurun_out/r2y/bench_default.json
