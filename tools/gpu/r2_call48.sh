#!/bin/bash
# round 2, GPU call 48: headline kernel at 5 CTAs x 192 threads / 3 CTAs x 288 threads per SM with the lifted operands re-lifted from the stash (RELIFT)
O=gpurun_out/r2av
mkdir -p $O
B="--workload cfg2 --steps 20 --no-cpu --no-e2e --no-others --no-sharded"
run() { # name, env...
  n=$1; shift
  env "$@" timeout 120 python bench.py $B > $O/$n.json 2> $O/$n.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/$n.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('$n', 'kernel_ms', round(r.get('kernel_ms'),4), 'frac', round(r.get('frac'),4))
except Exception as e: print('$n', 'ERR', e)
PY
}
run nt192_r68_relift PYSTL_B200_FUSED_REGS_BWD=68 PYSTL_B200_RELIFT=1
run nt288_r72_relift PYSTL_B200_FORCE_THREADS=288 PYSTL_B200_FUSED_REGS_BWD=72 PYSTL_B200_RELIFT=1
run nt192_r68 PYSTL_B200_FUSED_REGS_BWD=68
