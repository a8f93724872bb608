#!/bin/bash
# round 2, GPU call 11: pipe utilisation (XU / ALU / FMA / LSU) of the Until kernels from ncu (one launch each)
mkdir -p gpurun_out/r2k
M="sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed_pipe_xu.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread"
for w in cfg5s cfg5a cfg5d cfg5c cfg4 cfg5sb cfg5ab cfg3 cfg2; do
  ncu --metrics $M --clock-control none -k regex:pystl_fused -s 4 -c 1 --csv --log-file gpurun_out/r2k/pipes_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-e2e --no-others --no-sharded > gpurun_out/r2k/ncu_$w.log 2>&1
done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/r2k/pipes_*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>5 and r[0].isdigit()]
    d={r[-3]:r[-1] for r in rows}
    print(f.split('pipes_')[1][:-4], {k.split('__')[1][:34]:v for k,v in d.items()})
PY
