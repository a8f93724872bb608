#!/bin/bash
# round 2, GPU call 13: whole suite with the randomised Until adjoints and the NaN-flow corpus; register targets of the remaining Until adjoint kernels
mkdir -p gpurun_out/r2m
timeout 1500 python -m pytest tests -m gpu -q -n 4 > gpurun_out/r2m/pytest.log 2>&1
tail -5 gpurun_out/r2m/pytest.log
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg4 cfg5c cfg5sb cfg5ab; do
  timeout 300 python bench.py --workload $w $B > gpurun_out/r2m/$w.json 2> gpurun_out/r2m/$w.err
done
for v in "cfg5cb 2" "cfg5cb 3" "cfg5db 2" "cfg5db 3" "cfg5b 2" "cfg5b 3" "cfg5d 3"; do
  set -- $v
  PYSTL_B200_FUSED_MINB=$2 timeout 300 python bench.py --workload $1 $B > gpurun_out/r2m/$1_mb$2.json 2> gpurun_out/r2m/$1_mb$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
