#!/bin/bash
# round 2, GPU call 46: AGM Until forward with products of the Or-lift factors (one lg2 per four candidates); CT-STL mean test
O=gpurun_out/r2at
mkdir -p $O
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg5a cfg5ab; do
  timeout 300 python bench.py --workload $w $B > $O/bench_$w.json 2> $O/bench_$w.err
  python - <<PY
import json
d=json.loads(open('$O/bench_$w.json').read().strip().splitlines()[-1]); r=d['roofline']
print('$w', 'kernel_ms', r.get('kernel_ms'), 'frac', r.get('frac'))
PY
done
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_backward.py tests/test_gpu_parity.py tests/test_gpu_ctstl.py -x -q -m gpu -k "agm or ctstl or mean" > $O/pytest.log 2>&1 ) 2> $O/pytest.time
tail -3 $O/pytest.log; cat $O/pytest.time
