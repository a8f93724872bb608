# usage: WORKLOADS="cfg2" NTS="128 192 256 512" MINBS="0" FLAGS="--forward-only" bash tools/gpu/gpu_sweep.sh
mkdir -p gpurun_out
for w in ${WORKLOADS:-cfg2}; do for nt in ${NTS:-0}; do for mb in ${MINBS:-0}; do
if [ "$nt" != "0" ]; then export PYSTL_B200_FORCE_THREADS=$nt; else unset PYSTL_B200_FORCE_THREADS; fi
if [ "$mb" != "0" ]; then export PYSTL_B200_FUSED_MINB=$mb; else unset PYSTL_B200_FUSED_MINB; fi
python bench.py --workload $w ${FLAGS} --steps 10 --no-cpu --no-e2e > gpurun_out/sw_$w.json 2> gpurun_out/sw_$w.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/sw_$w.json')); r=d['roofline']
    s='$w nt=$nt minb=$mb %.3e samples/s | %s %.3f ms frac %.3f' % (d['value'], r['kernel'], r['kernel_ms'], r['frac'])
    for o in r.get('other_kernels', []): s+=' | %s %.3f ms frac %.3f' % (o['kernel'], o['kernel_ms'], o['frac'])
    print(s)
except Exception as e:
    print('$w nt=$nt minb=$mb FAILED', open('gpurun_out/sw_$w.err').read()[-300:])
PY
done; done; done
