#!/usr/bin/env python
"""Per-source-line view of one kernel of an ncu report: executed warp instructions per particle and
share of stall samples, joined with the SASS of the built library.

    python tools/ncu_hot.py gpurun_out/x.ncu-rep k_push_depositILi2ELb0 16777216 [launch_index]
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
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout.split("\n")
idx = [i for i, l in enumerate(raw) if l.startswith('"Kernel Name"')] + [len(raw)]
which = min(which, len(idx) - 2)
rows = list(csv.reader(io.StringIO("\n".join(raw[idx[which]:idx[which + 1]]))))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
base = int(data[0][ix["Address"]], 16)
samp = {int(r[ix["Address"]], 16) - base: (float(r[ix["# Samples"]] or 0), float(r[ix["Instructions Executed"]] or 0))
        for r in data}
os.makedirs("/tmp/ncu_hot", exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "nix_b200", "libnixb200.so")], cwd="/tmp/ncu_hot",
               capture_output=True)
cub = [f for f in os.listdir("/tmp/ncu_hot") if f.startswith("push_deposit") and f.endswith(".cubin")]
cubin = os.path.join("/tmp/ncu_hot", cub[0] if cub else sorted(os.listdir("/tmp/ncu_hot"))[0])
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur, infn = None, False
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter(), collections.Counter()])
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
    if m and int(m.group(1), 16) in samp:
        sa, ex = samp[int(m.group(1), 16)]
        a = agg[cur]
        a[0] += sa
        a[1] += ex
        a[2][m.group(3).split(".")[0]] += sa
        a[3][m.group(3).split(".")[0]] += ex
tot = sum(a[0] for a in agg.values())
totex = sum(a[1] for a in agg.values())
print(f"total {totex / npart:.1f} warp-instr/particle, {tot:.0f} samples")
ops = collections.Counter()
for a in agg.values():
    ops.update(a[3])
print("by opcode (warp-instr/particle):", " ".join(f"{o}:{c / npart:.1f}" for o, c in ops.most_common(24)))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{100 * a[0] / tot:5.1f}%  {a[1] / npart:6.2f}/p  {str(k):34s}",
          " ".join(f"{o}:{100 * c / tot:.1f}" for o, c in a[2].most_common(4)))
