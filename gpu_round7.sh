mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg4 cfg5s cfg5c; do WORKLOADS=$w bash gpu_sweep.sh; done
NTS="128 192 256 320 512" WORKLOADS="cfg4" bash gpu_sweep.sh
W=cfg4 TAG=r1k KERNEL=pystl_fused_forward FLAGS="" bash gpu_ncu.sh
W=cfg5s TAG=r1k KERNEL=pystl_fused_forward FLAGS="" bash gpu_ncu.sh
