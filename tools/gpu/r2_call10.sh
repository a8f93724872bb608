#!/bin/bash
# round 2, GPU call 10: D-GMSR Until forward fast path, AGM Until forward with the reciprocal table
mkdir -p gpurun_out/r2j
timeout 1500 python -m pytest tests -m gpu -q -x -n 4 > gpurun_out/r2j/pytest.log 2>&1
tail -3 gpurun_out/r2j/pytest.log
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg5a cfg5d cfg5ab cfg5db cfg4b cfg5cb; do
  timeout 300 python bench.py --workload $w $B > gpurun_out/r2j/$w.json 2> gpurun_out/r2j/$w.err
done
python tests/error_table.py --backend gpu --only dgmsr --out gpurun_out/r2j/err_dgmsr.md > gpurun_out/r2j/err_dgmsr.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), r.get('pipe_floor'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
grep -E "cfg5|cfg4" gpurun_out/r2j/err_dgmsr.md | head -20
