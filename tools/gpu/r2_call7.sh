#!/bin/bash
# round 2, GPU call 7: streamlined AGM Until adjoint, trimmed headline kernel (32-bit tile arithmetic, one-FFMA predicates,
# root n-ary adjoint folded into the forward loop, -0 initialised adjoints), smem-aware register target
mkdir -p gpurun_out/r2g
timeout 1500 python -m pytest tests -m gpu -q -x -n 4 > gpurun_out/r2g/pytest.log 2>&1
tail -3 gpurun_out/r2g/pytest.log
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg2 cfg3 cfg3d cfg1xb cfg5sb cfg5ab cfg4b cfg1x cfg5s; do
  timeout 300 python bench.py --workload $w $B > gpurun_out/r2g/$w.json 2> gpurun_out/r2g/$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2g/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), d['config']['plan']['threads'], d['config']['plan']['smem_bytes'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
