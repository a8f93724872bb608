#!/bin/bash
# round 2, GPU call 32 (element-per-thread assemble / scatter): assemble / scatter kernels, row sums: leaf-predicate + open-window tests, open-ended formulas through the public API
mkdir -p gpurun_out/r2af
timeout 900 python -m pytest tests/test_gpu_open_window.py tests/test_gpu_leaf_predicates.py tests/test_gpu_interface.py -m gpu -q > gpurun_out/r2af/pytest.log 2>&1
tail -4 gpurun_out/r2af/pytest.log
timeout 600 python tools/open_bench.py > gpurun_out/r2af/open_bench.jsonl 2> gpurun_out/r2af/open_bench.err
cut -c1-330 gpurun_out/r2af/open_bench.jsonl; tail -5 gpurun_out/r2af/open_bench.err
for c in 0 4; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2af/launches_case$c.csv \
  python tools/open_bench.py --only $c --iters 1 --warm 1 > gpurun_out/r2af/ncu_case$c.log 2>&1
done
python - <<'PY'
import csv,io
for c in (0,4):
    rows=[l for l in open(f'gpurun_out/r2af/launches_case{c}.csv') if l.startswith('"')]
    r=list(csv.DictReader(io.StringIO(''.join(rows))))
    print('case',c,len(r),'launches')
    for d in r[len(r)//2:]:
        print('  ',d['Kernel Name'][:90], d['Metric Value'])
PY
