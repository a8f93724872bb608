#!/bin/bash
# round 2, GPU call 17: per-line ncu capture of the cumulative value-and-gradient kernels (cfg5cb, cfg4b)
mkdir -p gpurun_out/r2q
export PYSTL_B200_NVRTC_FLAGS=-lineinfo
export PYSTL_B200_KEEP_SOURCE=1
for w in cfg5cb cfg4b; do
  python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-e2e --no-others --no-sharded > gpurun_out/r2q/plain_$w.json 2> gpurun_out/r2q/plain_$w.err
  ncu --set full --clock-control none --import-source on -k regex:pystl_fused_value_and_grad -s 3 -c 1 -f -o gpurun_out/r2q/prof_$w python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-e2e --no-others --no-sharded > gpurun_out/r2q/ncu_$w.log 2>&1
  ncu -i gpurun_out/r2q/prof_$w.ncu-rep --page source --csv --print-source sass > gpurun_out/r2q/prof_${w}_sass.csv 2> gpurun_out/r2q/sass_export_$w.err
done
cp pystl_b200/csrc/build/fused/*.cu gpurun_out/r2q/ 2>/dev/null
ls gpurun_out/r2q | head -30
