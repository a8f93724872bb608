mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg1xb cfg2 cfg3 cfg3d; do WORKLOADS=$w bash gpu_sweep.sh; done
