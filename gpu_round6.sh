mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg4 cfg5 cfg5s cfg5a cfg5d cfg5c; do WORKLOADS=$w bash gpu_sweep.sh; done
