for regs in ${REGS:-64 96 128 168}; do export PYSTL_B200_FUSED_REGS_BWD=$regs; echo "== bwd regs $regs"; bash gpu_sweep.sh; done
