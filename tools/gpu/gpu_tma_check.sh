python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for t in 1 0; do export PYSTL_B200_TMA_STORE=$t; echo "== TMA store $t"; for w in cfg2 cfg3 cfg3d; do WORKLOADS=$w bash tools/gpu/gpu_sweep.sh; done; done
