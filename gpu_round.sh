# usage: [WORKLOADS="cfg1 cfg2"] [NT="0 64 128"] bash gpu_round.sh [lib.so]  -- forward benches
[ -n "$1" ] && export PYSTL_B200_LIB=$1
for w in ${WORKLOADS:-cfg1 cfg2 cfg3 cfg3d cfg4 cfg5}; do for nt in ${NT:-0}; do
if [ "$nt" != "0" ]; then export PYSTL_B200_FORCE_THREADS=$nt; else unset PYSTL_B200_FORCE_THREADS; fi
python bench.py --workload $w --forward-only --steps 5 --no-cpu --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$w.json'))
p=d['config']['plan']
print('$w nt=$nt %.3e samples/s  %.3f ms  %.0f GB/s  frac %.3f  nt=%d lout=%d tiles=%d smem=%d' % (d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], p['threads'], p['tile_out'], p['n_tiles'], p['smem_bytes']))
" 2>/dev/null || echo "$w nt=$nt FAILED: $(tail -1 gpurun_out/bench_$w.err)"; done; done
