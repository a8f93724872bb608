#!/bin/bash
# round 2, GPU call 49 (the last two GPU-minutes): the wide Smooth value-and-gradient tile as the DEFAULT (288 threads, operands re-lifted) --
# cfg2 kernel line, then the Smooth GPU tests of the fused tier and of the backward pass, then as much of the whole suite as the budget allows
O=gpurun_out/r2aw
mkdir -p $O
timeout 20 python bench.py --workload cfg2 --steps 20 --no-cpu --no-e2e --no-others --no-sharded > $O/cfg2_default.json 2> $O/cfg2_default.err
python - <<PY
import json
try:
    d=json.loads(open('$O/cfg2_default.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('cfg2 default', 'kernel_ms', round(r.get('kernel_ms'),4), 'frac', round(r.get('frac'),4), 'value', d['value'])
except Exception as e: print('cfg2 default ERR', e)
PY
timeout 50 python -m pytest tests/test_gpu_fused.py tests/test_gpu_backward.py tests/test_gpu_parity.py -x -q -m gpu -k smooth > $O/pytest_smooth.log 2>&1
tail -2 $O/pytest_smooth.log
timeout 30 python -m pytest tests/test_gpu_interface.py tests/test_gpu_edges.py -x -q -m gpu > $O/pytest_iface.log 2>&1
tail -2 $O/pytest_iface.log
