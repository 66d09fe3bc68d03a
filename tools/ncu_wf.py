#!/usr/bin/env python
"""Shared-memory wavefronts per source line of one kernel of an ncu report (per 32 particles), with the
ideal count and the stall samples attributed to the line.

    python tools/ncu_wf.py gpurun_out/x.ncu-rep k_push_depositILi2ELb0 16777216 [launch_index] [lib.so]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

rep, fn, npart = sys.argv[1], sys.argv[2], float(sys.argv[3])
which = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[5] if len(sys.argv) > 5 else os.path.join(ROOT, "nix_b200", "libnixb200.so")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout.split("\n")
idx = [i for i, l in enumerate(raw) if l.startswith('"Kernel Name"')] + [len(raw)]
which = min(which, len(idx) - 2)
rows = list(csv.reader(io.StringIO("\n".join(raw[idx[which]:idx[which + 1]]))))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
base = int(data[0][ix["Address"]], 16)


def f(r, k):
    try:
        return float(r[ix[k]] or 0)
    except ValueError:
        return 0.0


info = {int(r[ix["Address"]], 16) - base: (f(r, "# Samples"), f(r, "Instructions Executed"), f(r, "L1 Wavefronts Shared"),
                                           f(r, "L1 Wavefronts Shared Ideal"), f(r, os.environ.get("STALL", "stall_long_sb"))) for r in data}
tmp = "/tmp/ncu_wf"
os.makedirs(tmp, exist_ok=True)
for x in os.listdir(tmp):
    os.remove(os.path.join(tmp, x))
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cub = [x for x in os.listdir(tmp) if x.endswith(".cubin") and fn.split("ILi")[0] in open(os.path.join(tmp, x), "rb").read().decode("latin1")]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub[0])], capture_output=True, text=True).stdout
cur, infn = None, False
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, collections.Counter(), 0.0, collections.Counter()])
for l in sass.split("\n"):
    if l.startswith(".text.") or l.startswith("//--------------------- .text"):
        infn = fn in l
    if not infn:
        continue
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        cur = m.group(1).split("/")[-1] + ":" + m.group(2)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m and int(m.group(1), 16) in info:
        sa, ex, wf, wfi, st = info[int(m.group(1), 16)]
        a = agg[cur]
        a[0] += sa; a[1] += ex; a[2] += wf; a[3] += wfi; a[5] += st
        if st:
            a[6][m.group(3)] += st
        if wf:
            a[4][m.group(3)] += wf
it = npart / 32
tw = sum(a[2] for a in agg.values())
ts = sum(a[0] for a in agg.values())
print(f"total {tw / it:.0f} shared wavefronts per 32 particles (ideal {sum(a[3] for a in agg.values()) / it:.0f}), "
      f"{sum(a[1] for a in agg.values()) / it:.0f} warp instr per 32 particles")
sk = {"wf": 2, "instr": 1, "samp": 0}[os.environ.get("SORT", "wf")]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][sk])[:int(os.environ.get('TOP', 30))]:
    if a[sk] == 0:
        break
    print(f"{a[2] / it:7.1f} wf (ideal {a[3] / it:6.1f})  {a[1] / it:6.1f} instr  {100 * a[0] / ts:4.1f}% samp  {str(k):34s}",
          " ".join(f"{o}:{c / it:.0f}" for o, c in a[4].most_common(3)))

if os.environ.get("STALL"):
    tot = sum(a[5] for a in agg.values())
    print(f"--- {os.environ['STALL']}: {tot:.0f} samples of {ts:.0f}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][5])[:int(os.environ.get('TOP', 30))]:
        if a[5] == 0:
            break
        print(f"{100 * a[5] / ts:5.1f}% of all samples  {a[1] / it:6.1f} instr  {str(k):34s}",
              " ".join(f"{o}:{100 * c / ts:.1f}" for o, c in a[6].most_common(3)))
