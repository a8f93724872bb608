#!/usr/bin/env python
"""Join an ncu SASS source page (per-address executed counts) with nvdisasm -gi line info.

    python tools/sass_by_line.py <ncu_sass.csv> <nvdisasm_gi.sass> <generated.cu> [--inner]

Aggregates executed warp instructions and stall samples by line of the generated kernel (the
outermost inlining frame), or with --inner by the innermost fused_rt.cuh line.
"""
import csv, re, sys, collections
ncu, sass, gen = sys.argv[1:4]
inner = '--inner' in sys.argv
rows = list(csv.reader(open(ncu)))
hdr = rows[1]; data = rows[2:]
iA, iE, iS = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
base = int(data[0][iA], 16)
cnt = {int(r[iA], 16) - base: (int(r[iE]), int(r[iS]), r[hdr.index('Source')]) for r in data}
outer = None; innermost = None; first = True
agg = collections.defaultdict(lambda: [0, 0, 0])
for line in open(sass):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        if first:
            innermost = (m.group(1).split('/')[-1], int(m.group(2))); first = False
        outer = (m.group(1).split('/')[-1], int(m.group(2)))
        if m.group(3): outer = (m.group(3).split('/')[-1], int(m.group(4)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', line)
    if m:
        off = int(m.group(1), 16)
        key = innermost if inner else outer
        if off in cnt:
            e, s, _ = cnt[off]
            a = agg[key]; a[0] += e; a[1] += s; a[2] += 1
        first = True
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
src = {}
try:
    src = {i + 1: l.rstrip() for i, l in enumerate(open(gen))}
except OSError:
    pass
print(f'total warp inst {tot}, samples {ts}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    text = src.get(k[1], '')[:90] if k[0].endswith('.cu') and not inner else ''
    print(f'{k[0]}:{k[1]:<5d} inst {a[0] / tot * 100:5.1f}%  stalls {a[1] / max(ts, 1) * 100:5.1f}%  static {a[2]:4d}  {text}')
