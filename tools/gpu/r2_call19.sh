#!/bin/bash
# round 2, GPU call 19: cumulative lockstep walk, FSET + FMA formulation with the tie in the main path
mkdir -p gpurun_out/r2s
timeout 1500 python -m pytest tests -m gpu -q -n 4 -k "cumulative or horizon or nan_values or random" > gpurun_out/r2s/pytest.log 2>&1
tail -3 gpurun_out/r2s/pytest.log
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg4b cfg5cb; do
  timeout 300 python bench.py --workload $w $B > gpurun_out/r2s/$w.json 2> gpurun_out/r2s/$w.err
done
PYSTL_B200_FUSED_MINB=1 timeout 300 python bench.py --workload cfg5cb $B > gpurun_out/r2s/cfg5cb_mb1.json 2> gpurun_out/r2s/cfg5cb_mb1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), d['config']['plan']['threads'], d['config']['plan']['smem_bytes'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
