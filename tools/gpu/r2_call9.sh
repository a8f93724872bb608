#!/bin/bash
# round 2, GPU call 9: cumulative Until adjoint on the lockstep walk, rare exact paths out of line
mkdir -p gpurun_out/r2i
timeout 1500 python -m pytest tests -m gpu -q -x -n 4 > gpurun_out/r2i/pytest.log 2>&1
tail -3 gpurun_out/r2i/pytest.log
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg2 cfg3 cfg4b cfg5cb cfg5sb cfg5s cfg4 cfg5c; do
  timeout 300 python bench.py --workload $w $B > gpurun_out/r2i/$w.json 2> gpurun_out/r2i/$w.err
done
PYSTL_B200_RELIFT=1 timeout 300 python bench.py --workload cfg2 $B > gpurun_out/r2i/cfg2_relift.json 2> gpurun_out/r2i/cfg2_relift.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), d['config']['plan']['threads'], d['config']['plan']['smem_bytes'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
