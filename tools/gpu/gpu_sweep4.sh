# value-and-gradient kernels: register target x CTA size
for regs in ${REGS:-80 96}; do export PYSTL_B200_FUSED_REGS_BWD=$regs; echo "== bwd regs $regs"; WORKLOADS="${WORKLOADS:-cfg2 cfg3 cfg3d}" NTS="${NTS:-0 128 160 192 224 256 320}" bash tools/gpu/gpu_sweep.sh; done
