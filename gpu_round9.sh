mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in cfg1xb cfg1x; do WORKLOADS=$w bash gpu_sweep.sh; done
python tools/bench_classic_grad.py 2>&1 | tail -3
