mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg1x cfg2 cfg3 cfg3d cfg4 cfg5; do WORKLOADS=$w bash gpu_sweep.sh; done
echo "== no merge"
export PYSTL_B200_NO_MERGE=1
for w in cfg2 cfg3 cfg5; do WORKLOADS=$w bash gpu_sweep.sh; done
