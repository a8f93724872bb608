#!/bin/bash
# round 2, last GPU call: the GPU suite as the driver runs it (serial, -x) on the final code, smoke(), the default bench line,
# launch list + one full ncu capture of cfg5 Smooth forward (the kernel that crossed 0.60 this session)
O=gpurun_out/r2final3
mkdir -p $O
( time timeout 1500 python -m pytest tests/ -x -q -m gpu > $O/pytest_serial.log 2>&1 ) 2> $O/pytest_serial.time
tail -3 $O/pytest_serial.log; cat $O/pytest_serial.time
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; head -c 200 $O/bench_default.json; echo; tail -2 $O/bench_default.err
W=cfg5s TAG=r2end KERNEL=pystl_fused_forward SKIP=5 bash tools/gpu/gpu_ncu.sh > $O/gpu_ncu.log 2>&1; tail -2 $O/gpu_ncu.log
