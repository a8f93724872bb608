#!/bin/bash
# round 2, GPU call 20: per-line ncu capture of the AGM and Smooth Until value-and-gradient kernels (cfg5ab, cfg5sb)
mkdir -p gpurun_out/r2t
export PYSTL_B200_NVRTC_FLAGS=-lineinfo
for w in cfg5ab cfg5sb; do
  python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-e2e --no-others --no-sharded > gpurun_out/r2t/plain_$w.json 2> gpurun_out/r2t/plain_$w.err
  ncu --set full --clock-control none --import-source on -k regex:pystl_fused_value_and_grad -s 3 -c 1 -f -o gpurun_out/r2t/prof_$w python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-e2e --no-others --no-sharded > gpurun_out/r2t/ncu_$w.log 2>&1
  ncu -i gpurun_out/r2t/prof_$w.ncu-rep --page source --csv --print-source sass > gpurun_out/r2t/prof_${w}_sass.csv 2> gpurun_out/r2t/sass_export_$w.err
done
ls gpurun_out/r2t | head
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
unset PYSTL_B200_NVRTC_FLAGS
timeout 600 python -m pytest tests -m gpu -q -n 4 -k "agm" > gpurun_out/r2t/pytest_agm.log 2>&1; tail -2 gpurun_out/r2t/pytest_agm.log
for w in cfg5a cfg5ab; do timeout 300 python bench.py --workload $w $B > gpurun_out/r2t/$w.json 2> gpurun_out/r2t/$w.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2t/$w.json').read().strip().splitlines()[-1]); print('$w', d['roofline']['kernel_ms'])"; done
