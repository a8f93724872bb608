#!/bin/bash
# round 2, GPU call 45: D-GMSR Until forward, warp-uniform case votes (PYSTL_DGMSR_SELECT=2) against evaluate-and-select (default)
O=gpurun_out/r2as
mkdir -p $O
B="--steps 10 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg5d; do
  timeout 300 python bench.py --workload $w $B > $O/bench_${w}_select.json 2> $O/bench_${w}_select.err
  PYSTL_B200_NVRTC_FLAGS=-DPYSTL_DGMSR_SELECT=2 timeout 300 python bench.py --workload $w $B > $O/bench_${w}_vote.json 2> $O/bench_${w}_vote.err
  python - <<PY
import json
for v in ('select','vote'):
    try:
        d=json.loads(open('$O/bench_${w}_'+v+'.json').read().strip().splitlines()[-1]); r=d['roofline']
        print('$w', v, 'kernel_ms', r.get('kernel_ms'), 'frac', r.get('frac'))
    except Exception as e: print('$w', v, 'ERR', e)
PY
done
tail -3 $O/*.err
PYSTL_B200_NVRTC_FLAGS=-DPYSTL_DGMSR_SELECT=2 timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -m gpu -k "dgmsr" 2>&1 | tail -3
