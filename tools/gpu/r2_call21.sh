#!/bin/bash
# round 2, GPU call 21: D-GMSR Until adjoint on the lockstep walk (closed forms per case); AGM forward reverted to one variant
mkdir -p gpurun_out/r2u
timeout 1500 python -m pytest tests -m gpu -q -n 4 > gpurun_out/r2u/pytest.log 2>&1
tail -3 gpurun_out/r2u/pytest.log
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg5db cfg5d cfg5a cfg5ab; do
  timeout 300 python bench.py --workload $w $B > gpurun_out/r2u/$w.json 2> gpurun_out/r2u/$w.err
done
python tests/error_table.py --backend gpu --only dgmsr --out gpurun_out/r2u/err_dgmsr.md > gpurun_out/r2u/err_dgmsr.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2u/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), d['config']['plan']['threads'], d['config']['plan']['smem_bytes'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
grep -E "cfg5|cfg4" gpurun_out/r2u/err_dgmsr.md | grep -v "n/a" | head -14
