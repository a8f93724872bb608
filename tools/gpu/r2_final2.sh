#!/bin/bash
# round 2, closing GPU call: the GPU suite exactly as the driver runs it (serial, -x), smoke(), the default bench line, cfg5 Smooth
# forward after the exponential planes (bench line + ncu pipes), ncu pipes of the CT-STL selection kernel
O=gpurun_out/r2final2
mkdir -p $O
( time timeout 1500 python -m pytest tests/ -x -q -m gpu --durations=10 > $O/pytest_serial.log 2>&1 ) 2> $O/pytest_serial.time
tail -4 $O/pytest_serial.log; cat $O/pytest_serial.time
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -5 $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; head -c 300 $O/bench_default.json; echo; tail -2 $O/bench_default.err
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
timeout 300 python bench.py --workload cfg5s $B > $O/bench_cfg5s.json 2> $O/bench_cfg5s.err
M="sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed_pipe_xu.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread"
ncu --metrics $M --clock-control none -k regex:pystl_fused -s 4 -c 1 --csv --log-file $O/pipes_cfg5s.csv python bench.py --workload cfg5s --steps 2 --warmup 1 --no-cpu --no-e2e --no-others --no-sharded > $O/ncu_pipes_cfg5s.log 2>&1
ncu --metrics $M --clock-control none -k regex:kth_window_fwd -s 30 -c 1 --csv --log-file $O/pipes_kth_window_fwd.csv python tools/ctstl_bench.py > $O/ncu_pipes_kth.log 2>&1
grep -h "pct_of_peak\|gpu__time" $O/pipes_cfg5s.csv $O/pipes_kth_window_fwd.csv | cut -d, -f5,13-15 | head -30
