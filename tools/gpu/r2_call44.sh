#!/bin/bash
# round 2, GPU call 44: D-GMSR Until forward with per-candidate case branches (against evaluate-and-select), selection kernel edge predicates
O=gpurun_out/r2ar
mkdir -p $O
B="--steps 10 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg5d cfg5db; do
  timeout 300 python bench.py --workload $w $B > $O/bench_${w}_branch.json 2> $O/bench_${w}_branch.err
  PYSTL_B200_NVRTC_FLAGS=-DPYSTL_DGMSR_SELECT=1 timeout 300 python bench.py --workload $w $B > $O/bench_${w}_select.json 2> $O/bench_${w}_select.err
  python - <<PY
import json
for v in ('branch','select'):
    try:
        d=json.loads(open('$O/bench_${w}_'+v+'.json').read().strip().splitlines()[-1]); r=d['roofline']
        print('$w', v, 'kernel_ms', r.get('kernel_ms'), 'frac', r.get('frac'))
    except Exception as e: print('$w', v, 'ERR', e)
PY
done
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_backward.py tests/test_gpu_parity.py -x -q -m gpu -k "dgmsr" > $O/pytest.log 2>&1 ) 2> $O/pytest.time
tail -3 $O/pytest.log; cat $O/pytest.time
timeout 300 python tools/ctstl_bench.py > $O/ctstl_bench.jsonl 2> $O/ctstl_bench.err; grep kth_window_fwd $O/ctstl_bench.jsonl | cut -c1-200
