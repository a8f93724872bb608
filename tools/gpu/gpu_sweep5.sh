WORKLOADS="cfg3 cfg3d" NTS="160 192 224 256 288" MINBS="2 3 4" bash tools/gpu/gpu_sweep.sh
WORKLOADS="cfg2" NTS="128 160 192" MINBS="3 4 5" bash tools/gpu/gpu_sweep.sh
