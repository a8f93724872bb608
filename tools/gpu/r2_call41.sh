#!/bin/bash
# round 2, GPU call 41: CT-STL kernels (select.cu): GPU tests, smoke, a first timing of the selection kernel
O=gpurun_out/r2ao
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_ctstl.py -x -q -m gpu --durations=5 > $O/pytest_ctstl.log 2>&1 ) 2> $O/pytest_ctstl.time
tail -25 $O/pytest_ctstl.log; cat $O/pytest_ctstl.time
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -4 $O/smoke.log
timeout 300 python tools/ctstl_bench.py > $O/ctstl_bench.jsonl 2> $O/ctstl_bench.err; cat $O/ctstl_bench.jsonl; tail -3 $O/ctstl_bench.err
