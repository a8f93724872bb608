#!/bin/bash
# round 2, GPU call 42: CT-STL GPU tests after the -0 / +0 key fix
O=gpurun_out/r2ap
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_ctstl.py -q -m gpu --durations=5 > $O/pytest_ctstl.log 2>&1 ) 2> $O/pytest_ctstl.time
tail -40 $O/pytest_ctstl.log | cut -c1-400; cat $O/pytest_ctstl.time
