#!/bin/bash
# round 2, GPU call 38: default bench with the open-ended lines
mkdir -p gpurun_out/r2al
( time timeout 900 python bench.py > gpurun_out/r2al/bench_default.json 2> gpurun_out/r2al/bench_default.err ) 2> gpurun_out/r2al/bench.time
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2al/bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'])
for r in d['open_ended']: print(r)
PY
cat gpurun_out/r2al/bench.time; tail -3 gpurun_out/r2al/bench_default.err
