#!/bin/bash
# round 2, GPU call 2: whole GPU suite with every tolerance multiplier capped at 1 (what really needs more?), parity log,
# smoke(), L2-resident vs HBM-resident cfg2 (is the headline kernel latency- or issue-bound?), ncu capture with line info
mkdir -p gpurun_out/r2b
export PYSTL_PARITY_LOG=$PWD/gpurun_out/r2b/parity_log.tsv
PYSTL_PARITY_SCALE_CAP=1 timeout 1500 python -m pytest tests -m gpu -q -n 4 > gpurun_out/r2b/pytest_cap1.log 2>&1
unset PYSTL_PARITY_LOG
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2b/smoke.log 2>&1
for b in 1024 2048 4096 65536; do
  python bench.py --workload cfg2 --batch $b --steps 20 --no-cpu --no-e2e > gpurun_out/r2b/cfg2_b$b.json 2> gpurun_out/r2b/cfg2_b$b.err
done
for nt in 288 544; do
  PYSTL_B200_FORCE_THREADS=$nt PYSTL_B200_FUSED_MINB=1 python bench.py --workload cfg2 --steps 20 --no-cpu --no-e2e > gpurun_out/r2b/cfg2_nt$nt.json 2> gpurun_out/r2b/cfg2_nt$nt.err
done
export PYSTL_B200_NVRTC_FLAGS=-lineinfo
python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2b/ncu_plain.json 2> gpurun_out/r2b/ncu_plain.err
ncu --set full --clock-control none --import-source on -k regex:pystl_fused_value_and_grad -s 8 -c 1 -f -o gpurun_out/r2b/prof_cfg2_vg python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2b/ncu_full.log 2>&1
tail -5 gpurun_out/r2b/pytest_cap1.log
cat gpurun_out/r2b/smoke.log | tail -6
