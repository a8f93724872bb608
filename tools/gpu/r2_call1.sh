#!/bin/bash
# round 2, GPU call 1: measured error table (default and precise transcendentals), GPU test suite, baseline bench of every workload
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt
python tests/error_table.py --backend gpu --out gpurun_out/r2a/err_default.md > gpurun_out/r2a/err_default.log 2>&1
PYSTL_B200_NVRTC_FLAGS=-DPYSTL_PRECISE_MATH=15 python tests/error_table.py --backend gpu --out gpurun_out/r2a/err_precise15.md > gpurun_out/r2a/err_precise15.log 2>&1
for m in 1 2 4 8; do
PYSTL_B200_NVRTC_FLAGS=-DPYSTL_PRECISE_MATH=$m python tests/error_table.py --backend gpu --only dgmsr --out gpurun_out/r2a/err_precise$m.md > gpurun_out/r2a/err_precise$m.log 2>&1
done
timeout 1200 python -m pytest tests -m gpu -x -q -n 4 > gpurun_out/r2a/pytest.log 2>&1
for w in cfg1 cfg1x cfg1xb cfg2 cfg3 cfg3d cfg4 cfg4b cfg5 cfg5s cfg5a cfg5d cfg5c cfg5b cfg5sb cfg5ab; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu --no-e2e > gpurun_out/r2a/bench_$w.json 2> gpurun_out/r2a/bench_$w.err
done
tail -3 gpurun_out/r2a/pytest.log
