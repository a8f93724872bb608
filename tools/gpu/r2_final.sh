#!/bin/bash
# round 2, final GPU call: whole GPU suite with the parity log, smoke(), error table, the default bench line, one kernel line per
# workload, and the ncu evidence of the default bench (launch list, full capture of the headline kernel, pipe utilisation)
O=gpurun_out/r2final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,driver_version --format=csv > $O/smi.txt
export PYSTL_PARITY_LOG=$PWD/$O/parity_log.tsv
timeout 1500 python -m pytest tests -m gpu -q -n 4 > $O/pytest.log 2>&1
unset PYSTL_PARITY_LOG
tail -3 $O/pytest.log
python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke.log 2>&1
tail -6 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
B="--steps 20 --no-cpu --no-e2e --no-others --no-sharded"
for w in cfg1 cfg1x cfg1xb cfg2 cfg3 cfg3d cfg4 cfg4b cfg5 cfg5s cfg5a cfg5d cfg5c cfg5b cfg5sb cfg5ab cfg5db cfg5cb; do
  timeout 300 python bench.py --workload $w $B > $O/bench_$w.json 2> $O/bench_$w.err
done
python tests/error_table.py --backend gpu --out $O/err_table.md > $O/err_table.log 2>&1
# ncu: launch list of the default bench command (bounded), full capture of the headline kernel, pipes of every workload
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-others --no-sharded > $O/ncu_launches.log 2>&1
PYSTL_B200_NVRTC_FLAGS=-lineinfo ncu --set full --clock-control none --import-source on -k regex:pystl_fused_value_and_grad -s 6 -c 1 -f -o $O/prof_cfg2_vg python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu --no-e2e --no-others --no-sharded > $O/ncu_full.log 2>&1
M="sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed_pipe_xu.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread"
for w in cfg1x cfg1xb cfg2 cfg3 cfg3d cfg4 cfg4b cfg5 cfg5s cfg5a cfg5d cfg5c cfg5b cfg5sb cfg5ab cfg5db cfg5cb; do
  ncu --metrics $M --clock-control none -k regex:pystl_fused -s 4 -c 1 --csv --log-file $O/pipes_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-e2e --no-others --no-sharded > $O/ncu_pipes_$w.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2final/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline',{})
        print(f.split('/')[-1], 'value', d.get('value'), 'step', round(d.get('ms_per_step',0),4), 'kernel', round(r.get('kernel_ms',0),4), 'frac', round(r.get('frac',0),4), 'e2e', (d.get('e2e') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
