mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg3; do WORKLOADS=$w bash gpu_sweep.sh; done
python bench.py --no-cpu > gpurun_out/r8_default.json 2> gpurun_out/r8_default.err; tail -c 500 gpurun_out/r8_default.json
