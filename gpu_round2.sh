# usage: bash gpu_round2.sh  -- fused tier: tests, A/B bench, ncu
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_fused.log
cat gpurun_out/pytest_fused.log | tail -5
for tier in fused interp; do for w in cfg1 cfg2 cfg3 cfg3d; do
PYSTL_B200_TIER=$tier python bench.py --workload $w --forward-only --steps 10 --no-cpu --no-e2e > gpurun_out/r2_${tier}_$w.json 2> gpurun_out/r2_${tier}_$w.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_${tier}_$w.json')); r=d['roofline']
    print('$tier $w %.3e samples/s  %.3f ms  frac %.3f' % (d['value'], r['kernel_ms'], r['frac']))
except Exception as e:
    print('$tier $w FAILED', open('gpurun_out/r2_${tier}_$w.err').read()[-300:])
PY
done; done
