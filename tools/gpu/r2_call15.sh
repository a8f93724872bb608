#!/bin/bash
# round 2, GPU call 15: cumulative Until adjoint on the lockstep walk, threshold formulation (one compare + predicated add / count per candidate)
mkdir -p gpurun_out/r2o
timeout 1500 python -m pytest tests -m gpu -q -n 4 -k "cumulative or horizon or nan_values or random" > gpurun_out/r2o/pytest.log 2>&1
tail -3 gpurun_out/r2o/pytest.log
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg4b cfg5cb; do
  timeout 300 python bench.py --workload $w $B > gpurun_out/r2o/$w.json 2> gpurun_out/r2o/$w.err
  PYSTL_B200_OLD_CUM=1 timeout 300 python bench.py --workload $w $B > gpurun_out/r2o/${w}_old.json 2> gpurun_out/r2o/${w}_old.err
done
for v in "cfg5cb 2" "cfg5cb 3" "cfg4b 1"; do
  set -- $v
  PYSTL_B200_FUSED_MINB=$2 timeout 300 python bench.py --workload $1 $B > gpurun_out/r2o/$1_mb$2.json 2> gpurun_out/r2o/$1_mb$2.err
done
PYSTL_B200_FORCE_THREADS=256 timeout 300 python bench.py --workload cfg4b $B > gpurun_out/r2o/cfg4b_nt256.json 2> gpurun_out/r2o/cfg4b_nt256.err
PYSTL_B200_FORCE_THREADS=320 timeout 300 python bench.py --workload cfg4b $B > gpurun_out/r2o/cfg4b_nt320.json 2> gpurun_out/r2o/cfg4b_nt320.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2o/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), d['config']['plan']['threads'], d['config']['plan']['smem_bytes'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
