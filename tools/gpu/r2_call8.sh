#!/bin/bash
# round 2, GPU call 8: headline-kernel experiments (re-lift instead of keeping the lifted operands, prefetch distance, merged
# live blocks) and a per-line ncu capture of the cfg3 AGM value-and-gradient kernel
mkdir -p gpurun_out/r2h
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --workload cfg2 $B > gpurun_out/r2h/cfg2_$tag.json 2> gpurun_out/r2h/cfg2_$tag.err; }
run base X=1
run relift PYSTL_B200_RELIFT=1
run relift_mb5 PYSTL_B200_RELIFT=1 PYSTL_B200_FUSED_MINB=5
run pf0 PYSTL_B200_PREFETCH=0
run pf300 PYSTL_B200_PREFETCH=300
run pf1500 PYSTL_B200_PREFETCH=1500
run merge0 PYSTL_B200_MERGE=0
run r72 PYSTL_B200_NVRTC_FLAGS=-maxrregcount=72
run r96 PYSTL_B200_FUSED_MINB=3
export PYSTL_B200_NVRTC_FLAGS=-lineinfo
python bench.py --workload cfg3 --steps 2 --warmup 1 --no-cpu --no-e2e --no-others --no-sharded > gpurun_out/r2h/ncu_plain.json 2> gpurun_out/r2h/ncu_plain.err
ncu --set full --clock-control none --import-source on -k regex:pystl_fused_value_and_grad -s 4 -c 1 -f -o gpurun_out/r2h/prof_cfg3_vg python bench.py --workload cfg3 --steps 2 --warmup 1 --no-cpu --no-e2e --no-others --no-sharded > gpurun_out/r2h/ncu_full.log 2>&1
ncu -i gpurun_out/r2h/prof_cfg3_vg.ncu-rep --page source --csv --print-source sass > gpurun_out/r2h/prof_cfg3_vg_sass.csv 2> gpurun_out/r2h/sass_export.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h/cfg2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms'],4), 'frac', round(r['frac'],4))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
