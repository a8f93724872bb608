#!/bin/bash
# round 2, GPU call 6: whole GPU suite after the shared-memory re-layout (staging aliases everything behind the reduction
# scratch; lockstep until adjoints of smooth / agm in the dead window scratch), bench of the affected workloads
mkdir -p gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -q -x -n 4 > gpurun_out/r2f/pytest.log 2>&1
tail -3 gpurun_out/r2f/pytest.log
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg2 cfg3 cfg1xb cfg5 cfg5s cfg5b cfg5sb cfg5ab cfg5cb cfg4b; do
  timeout 300 python bench.py --workload $w $B > gpurun_out/r2f/$w.json 2> gpurun_out/r2f/$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), d['config']['plan']['threads'], d['config']['plan']['smem_bytes'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
