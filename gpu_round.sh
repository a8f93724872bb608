# usage: [WORKLOADS="cfg1 cfg2"] [NT="0 64 128"] [FLAGS="--forward-only"] bash gpu_round.sh [lib.so]
[ -n "$1" ] && export PYSTL_B200_LIB=$1
for w in ${WORKLOADS:-cfg1 cfg2 cfg3 cfg3d cfg4 cfg5}; do for nt in ${NT:-0}; do
if [ "$nt" != "0" ]; then export PYSTL_B200_FORCE_THREADS=$nt; else unset PYSTL_B200_FORCE_THREADS; fi
python bench.py --workload $w ${FLAGS} --steps 5 --no-cpu --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$w.json'))
p=d['config']['plan']; r=d['roofline']; o=r.get('other_kernel')
s='$w nt=$nt %.3e samples/s %.3f ms/step | %s %.3f ms frac %.3f' % (d['value'], d['ms_per_step'], r['kernel'], r['kernel_ms'], r['frac'])
if o: s+=' | %s %.3f ms frac %.3f' % (o['kernel'], o['kernel_ms'], o['frac'])
print(s, '| fwd plan nt=%d lout=%d tiles=%d smem=%d' % (p['threads'], p['tile_out'], p['n_tiles'], p['smem_bytes']))
" 2>/dev/null || echo "$w nt=$nt FAILED: $(tail -1 gpurun_out/bench_$w.err)"; done; done
