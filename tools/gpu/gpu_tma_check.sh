python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for t in 1 0; do export PYSTL_B200_TMA=$t; echo "== TMA $t"; for w in cfg1x cfg1xb cfg2 cfg3 cfg3d; do WORKLOADS=$w bash tools/gpu/gpu_sweep.sh; done; done
