for regs in 48 56 64; do export PYSTL_B200_FUSED_REGS=$regs; echo "== fwd regs $regs"; FLAGS="--forward-only" WORKLOADS="cfg1x cfg2 cfg3 cfg3d cfg4 cfg5 cfg5s cfg5c" bash tools/gpu/gpu_sweep.sh; done
