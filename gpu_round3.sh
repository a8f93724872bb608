mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -8
for w in cfg2 cfg3 cfg3d; do WORKLOADS=$w bash gpu_sweep.sh; done
PYSTL_B200_TIER=interp WORKLOADS=cfg2 bash gpu_sweep.sh
