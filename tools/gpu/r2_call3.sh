#!/bin/bash
# round 2, GPU call 3 (after the container was re-created): GPU suite, smoke(), default bench line, kernel lines of every workload
mkdir -p gpurun_out/r2c
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c/smi.txt
export PYSTL_PARITY_LOG=$PWD/gpurun_out/r2c/parity_log.tsv
timeout 1500 python -m pytest tests -m gpu -q -n 4 > gpurun_out/r2c/pytest.log 2>&1
unset PYSTL_PARITY_LOG
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2c/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/r2c/bench_default.json 2> gpurun_out/r2c/bench_default.err
for w in cfg1 cfg1x cfg1xb cfg2 cfg3 cfg3d cfg4 cfg4b cfg5 cfg5s cfg5a cfg5d cfg5c cfg5b cfg5sb cfg5ab cfg5db cfg5cb; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu --no-e2e --no-others --no-sharded > gpurun_out/r2c/bench_$w.json 2> gpurun_out/r2c/bench_$w.err
done
python tests/error_table.py --backend gpu --out gpurun_out/r2c/err_default.md > gpurun_out/r2c/err_default.log 2>&1
tail -3 gpurun_out/r2c/pytest.log
tail -6 gpurun_out/r2c/smoke.log
