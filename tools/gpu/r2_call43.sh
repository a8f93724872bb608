#!/bin/bash
# round 2, GPU call 43: the selection kernel with float compares (FSET on the ALU pipe + FADD on the FMA pipe): GPU tests, timing
O=gpurun_out/r2aq
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_ctstl.py -q -m gpu > $O/pytest_ctstl.log 2>&1 ) 2> $O/pytest_ctstl.time
tail -30 $O/pytest_ctstl.log | cut -c1-300; cat $O/pytest_ctstl.time
timeout 300 python tools/ctstl_bench.py > $O/ctstl_bench.jsonl 2> $O/ctstl_bench.err; grep kth_window_fwd $O/ctstl_bench.jsonl | cut -c1-260; tail -3 $O/ctstl_bench.err
