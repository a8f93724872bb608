python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for t in 1 0; do export PYSTL_B200_TMA=$t; echo "== TMA load $t"; for w in cfg1x cfg2 cfg4 cfg5 cfg5s cfg5c; do WORKLOADS=$w bash tools/gpu/gpu_sweep.sh; done; done
