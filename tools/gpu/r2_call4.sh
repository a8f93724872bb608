#!/bin/bash
# round 2, GPU call 4: CTA-size sweep of the headline kernel and an ncu capture with line info (per-line instruction counts)
mkdir -p gpurun_out/r2d
for nt in 192 256 288 384 512; do
  PYSTL_B200_FORCE_THREADS=$nt python bench.py --workload cfg2 --steps 20 --no-cpu --no-e2e --no-others --no-sharded > gpurun_out/r2d/cfg2_nt$nt.json 2> gpurun_out/r2d/cfg2_nt$nt.err
done
export PYSTL_B200_NVRTC_FLAGS=-lineinfo
python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu --no-e2e --no-others --no-sharded > gpurun_out/r2d/ncu_plain.json 2> gpurun_out/r2d/ncu_plain.err
ncu --set full --clock-control none --import-source on -k regex:pystl_fused_value_and_grad -s 4 -c 1 -f -o gpurun_out/r2d/prof_cfg2_vg python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu --no-e2e --no-others --no-sharded > gpurun_out/r2d/ncu_full.log 2>&1
ncu -i gpurun_out/r2d/prof_cfg2_vg.ncu-rep --page source --csv --print-source sass > gpurun_out/r2d/prof_cfg2_vg_sass.csv 2> gpurun_out/r2d/sass_export.err
cp pystl_b200/csrc/build/fused/*.cu gpurun_out/r2d/ 2>/dev/null
tail -3 gpurun_out/r2d/ncu_full.log
for nt in 192 256 288 384 512; do tail -c 400 gpurun_out/r2d/cfg2_nt$nt.json | head -c 300; echo; done
