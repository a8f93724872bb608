# SKIP = matching launches to skip: bench.py launches 5 steps (with the fused mean) first, then the
# kernel-only launches its roofline is computed from: SKIP=5 (forward workloads) / SKIP=8 (value-and-gradient)
# captures one of those.
# usage: W=cfg2 TAG=r1f FLAGS="--forward-only" KERNEL=pystl_fused bash tools/gpu/gpu_ncu.sh   -- launch list + one full capture
mkdir -p gpurun_out
export PYSTL_B200_NVRTC_FLAGS=-lineinfo
W=${W:-cfg2}; TAG=${TAG:-x}; K=${KERNEL:-pystl_fused_value_and_grad}
python bench.py --workload $W ${FLAGS} --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_plain_$TAG.json 2> gpurun_out/ncu_plain_$TAG.err || { tail -3 gpurun_out/ncu_plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${W}_$TAG.csv python bench.py --workload $W ${FLAGS} --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-3} -c 1 -f -o gpurun_out/prof_${W}_$TAG python bench.py --workload $W ${FLAGS} --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_f_$TAG.log 2>&1
ls -la gpurun_out/prof_${W}_$TAG.ncu-rep
