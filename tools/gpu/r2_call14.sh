#!/bin/bash
# round 2, GPU call 14: is the headline kernel sensitive to where its data comes from?  (L2-resident batches vs HBM)
mkdir -p gpurun_out/r2n
B="--steps 50 --no-cpu --no-e2e --no-others --no-sharded"
for b in 592 1184 2368 4736 65536; do
  timeout 300 python bench.py --workload cfg2 --batch $b $B > gpurun_out/r2n/cfg2_b$b.json 2> gpurun_out/r2n/cfg2_b$b.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2n/*.json'), key=lambda s:int(s.split('_b')[1][:-5])):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'kernel_ms', round(r['kernel_ms'],5), 'ns/trace', round(r['kernel_ms']*1e6/r['traces'],2), 'frac', round(r['frac'],4))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
