#!/bin/bash
# round 2, GPU call 30: where the time of the open-ended formulas goes (kernel list under ncu), after Assemble
mkdir -p gpurun_out/r2ad
timeout 600 python tools/open_bench.py > gpurun_out/r2ad/open_bench.jsonl 2> gpurun_out/r2ad/open_bench.err
cut -c1-330 gpurun_out/r2ad/open_bench.jsonl; tail -5 gpurun_out/r2ad/open_bench.err
for c in 0 4; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ad/launches_case$c.csv \
  python tools/open_bench.py --only $c --iters 1 --warm 1 > gpurun_out/r2ad/ncu_case$c.log 2>&1
done
python - <<'PY'
import csv,io
for c in (0,4):
    rows=[l for l in open(f'gpurun_out/r2ad/launches_case{c}.csv') if l.startswith('"')]
    r=list(csv.DictReader(io.StringIO(''.join(rows))))
    print('case',c,len(r),'launches')
    for d in r[len(r)//2:]:
        print('  ',d['Kernel Name'][:90], d['Metric Value'])
PY
