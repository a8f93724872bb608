#!/bin/bash
# round 2, GPU call 16: cumulative adjoint planes in dead forward scratch (cfg5: 141 -> 108 KB, 2 CTAs per SM); whole suite
mkdir -p gpurun_out/r2p
timeout 1500 python -m pytest tests -m gpu -q -n 4 > gpurun_out/r2p/pytest.log 2>&1
tail -3 gpurun_out/r2p/pytest.log
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg4b cfg5cb cfg5c cfg4; do
  timeout 300 python bench.py --workload $w $B > gpurun_out/r2p/$w.json 2> gpurun_out/r2p/$w.err
done
for v in "cfg5cb 1" "cfg5cb 2"; do
  set -- $v
  PYSTL_B200_FUSED_MINB=$2 timeout 300 python bench.py --workload $1 $B > gpurun_out/r2p/$1_mb$2.json 2> gpurun_out/r2p/$1_mb$2.err
done
PYSTL_B200_FORCE_THREADS=256 timeout 300 python bench.py --workload cfg4b $B > gpurun_out/r2p/cfg4b_nt256.json 2> gpurun_out/r2p/cfg4b_nt256.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), d['config']['plan']['threads'], d['config']['plan']['smem_bytes'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
