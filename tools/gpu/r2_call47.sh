#!/bin/bash
# round 2, GPU call 47: batching rules of the CT-STL ops (vmap over traces, vmap(grad))
O=gpurun_out/r2au
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ctstl.py -x -q -m gpu -k "vmap or mean" > $O/pytest.log 2>&1; tail -30 $O/pytest.log | cut -c1-250
