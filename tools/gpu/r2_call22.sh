#!/bin/bash
# round 2, GPU call 22: two reciprocals from one MUFU in the smooth Until walks (XU-bound kernels)
mkdir -p gpurun_out/r2v
timeout 1500 python -m pytest tests -m gpu -q -n 4 -k "smooth" > gpurun_out/r2v/pytest.log 2>&1
tail -3 gpurun_out/r2v/pytest.log
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg5s cfg5sb cfg2; do
  timeout 300 python bench.py --workload $w $B > gpurun_out/r2v/$w.json 2> gpurun_out/r2v/$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2v/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
