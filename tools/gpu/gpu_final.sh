# Final evidence run: tests, smoke, default bench (+ reference arm), every workload.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_pytest.log; cat gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -6 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; tail -c 600 gpurun_out/final_bench_default.json; echo
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -c 300 gpurun_out/final_bench_reference.json; echo
for w in cfg1 cfg1x cfg1xb cfg2 cfg3 cfg3d cfg4 cfg5 cfg5s cfg5c cfg5a cfg5d; do
  python bench.py --workload $w --steps 20 --no-cpu --no-e2e > gpurun_out/final_$w.json 2> gpurun_out/final_$w.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/final_$w.json')); r=d['roofline']
    s='$w %.3e samples/s | %s %.3f ms frac %.3f' % (d['value'], r['kernel'], r['kernel_ms'], r['frac'])
    for o in r.get('other_kernels', []): s+=' | %s %.3f ms frac %.3f' % (o['kernel'], o['kernel_ms'], o['frac'])
    print(s)
except Exception as e:
    print('$w FAILED', open('gpurun_out/final_$w.err').read()[-300:])
PY
done
