#!/bin/bash
# round 2, GPU call 39: open-ended formulas through the public API after removing a host-side stall (scalar H2D copy in the root-leaf mean + gradient)
mkdir -p gpurun_out/r2am
for i in 1 2; do
timeout 600 python tools/open_bench.py > gpurun_out/r2am/open_bench_$i.jsonl 2> gpurun_out/r2am/open_bench_$i.err
python - <<PY
import json
for l in open('gpurun_out/r2am/open_bench_$i.jsonl'):
    r=json.loads(l); print(r['semantics'], r['formula'], r['values_ms'], r['mean_grad_ms'])
PY
done
