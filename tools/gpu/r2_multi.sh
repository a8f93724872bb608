#!/bin/bash
# round 2: the driver's multi-GPU launch line on N GPUs of one node (N = number of visible GPUs)
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out/r2multi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2multi/bench_n$N.json 2> gpurun_out/r2multi/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2multi/ref_n$N.json 2> gpurun_out/r2multi/ref_n$N.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2multi/*.json')):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], {k: d.get(k) for k in ('impl','n_gpus','value','ms_per_step','scaling')}, 'e2e', (d.get('e2e') or {}).get('value'), 'sharded', {k: round(v['ms_per_step'],3) for k,v in ((d.get('cfg5_sharded') or {}).get('semantics') or {}).items()})
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
