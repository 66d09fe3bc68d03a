#!/usr/bin/env python
"""Join an ncu SASS source page (csv) with nvdisasm line info: executed instructions per source line.
    python tools/ncu_lines.py /tmp/src.csv /tmp/pd.sass k_push_depositILi2ELb0 nparticles"""
import collections
import csv
import re
import sys

srccsv, sass, fn, npart = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][ix["Address"]], 16)
execd = {}
for r in data:
    off = int(r[ix["Address"]], 16) - base
    execd[off] = (float(r[ix["Instructions Executed"]] or 0), float(r[ix["# Samples"]] or 0),
                  float(r[ix["L1 Wavefronts Shared"]] or 0), r[ix["Source"]].strip())
infn = False
cur = None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, collections.Counter()])
for l in open(sass):
    if l.startswith('.text.') or l.startswith('//--------------------- .text'):
        infn = fn in l
    if not infn:
        continue
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        cur = m.group(1).split('/')[-1] + ':' + m.group(2)
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
    if m:
        off = int(m.group(1), 16)
        if off in execd:
            e = execd[off]
            a = agg[cur]
            a[0] += e[0]
            a[1] += e[1]
            a[2] += e[2]
            a[3][m.group(3).split('.')[0] + ('.MOV' if '.MOV' in m.group(3) else '')] += e[0]
tot = sum(a[0] for a in agg.values())
ts = sum(a[1] for a in agg.values())
print(f"total {tot / npart:.1f} instr/particle")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    ops = ' '.join(f"{o}:{c / npart:.1f}" for o, c in a[3].most_common(4))
    print(f"{a[0] / npart:6.2f}/p  samp {100 * a[1] / ts:4.1f}%  wf {a[2] / npart:5.2f}  {k:28s} {ops}")
