#!/bin/bash
# round 2, GPU call 40: Smooth Until forward with exponential planes (PYSTL_B200_UNTIL_EXP): cfg5s with and without, fused GPU tests
O=gpurun_out/r2an
mkdir -p $O
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for e in 1 0 1 0; do
  PYSTL_B200_UNTIL_EXP=$e timeout 300 python bench.py --workload cfg5s $B > $O/bench_cfg5s_exp$e.json 2> $O/bench_cfg5s_exp$e.err
  python - <<PY
import json
d=json.loads(open('$O/bench_cfg5s_exp$e.json').read().strip().splitlines()[-1]); r=d['roofline']
print('exp$e', 'kernel_ms', r.get('kernel_ms'), 'frac', r.get('frac'))
PY
done
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -x -q -m gpu > $O/pytest.log 2>&1 ) 2> $O/pytest.time
tail -3 $O/pytest.log; cat $O/pytest.time
