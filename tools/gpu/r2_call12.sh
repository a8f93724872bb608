#!/bin/bash
# round 2, GPU call 12: one barrier per lockstep step (read-modify-write hand-off), raised limits; register targets of the Until adjoint kernels
mkdir -p gpurun_out/r2l
timeout 1500 python -m pytest tests -m gpu -q -x -n 4 > gpurun_out/r2l/pytest.log 2>&1
tail -3 gpurun_out/r2l/pytest.log
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg5sb cfg5ab cfg1 cfg2 cfg5d; do
  timeout 300 python bench.py --workload $w $B > gpurun_out/r2l/$w.json 2> gpurun_out/r2l/$w.err
done
PYSTL_B200_FUSED_MINB=2 timeout 300 python bench.py --workload cfg5sb $B > gpurun_out/r2l/cfg5sb_mb2.json 2> gpurun_out/r2l/cfg5sb_mb2.err
PYSTL_B200_FUSED_MINB=3 timeout 300 python bench.py --workload cfg5ab $B > gpurun_out/r2l/cfg5ab_mb3.json 2> gpurun_out/r2l/cfg5ab_mb3.err
PYSTL_B200_FUSED_MINB=2 timeout 300 python bench.py --workload cfg5d $B > gpurun_out/r2l/cfg5d_mb2.json 2> gpurun_out/r2l/cfg5d_mb2.err
PYSTL_B200_FUSED_MINB=4 timeout 300 python bench.py --workload cfg5d $B > gpurun_out/r2l/cfg5d_mb4.json 2> gpurun_out/r2l/cfg5d_mb4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
