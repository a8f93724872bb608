#!/bin/bash
# round 2, GPU call 5: the lockstep smooth Until adjoint on hardware (parity + cfg5sb), occupancy sweep of the headline kernel
mkdir -p gpurun_out/r2e
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_backward.py -m gpu -q -x -n 4 -k "smooth" > gpurun_out/r2e/pytest_smooth.log 2>&1
tail -3 gpurun_out/r2e/pytest_smooth.log
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
timeout 300 python bench.py --workload cfg5sb $B > gpurun_out/r2e/cfg5sb.json 2> gpurun_out/r2e/cfg5sb.err
for v in "192 4" "192 5" "160 5" "160 6" "128 6" "128 7" "128 8" "224 4"; do
  set -- $v
  PYSTL_B200_FORCE_THREADS=$1 PYSTL_B200_FUSED_MINB=$2 timeout 300 python bench.py --workload cfg2 $B > gpurun_out/r2e/cfg2_nt$1_mb$2.json 2> gpurun_out/r2e/cfg2_nt$1_mb$2.err
done
PYSTL_B200_TMA=2 timeout 300 python bench.py --workload cfg2 $B > gpurun_out/r2e/cfg2_tma2.json 2> gpurun_out/r2e/cfg2_tma2.err
PYSTL_B200_TMA=2 PYSTL_B200_FUSED_MINB=5 timeout 300 python bench.py --workload cfg2 $B > gpurun_out/r2e/cfg2_tma2_mb5.json 2> gpurun_out/r2e/cfg2_tma2_mb5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2e/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4), d['config']['plan']['threads'], d['config']['plan']['smem_bytes'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
