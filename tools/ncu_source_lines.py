#!/usr/bin/env python
"""Per-source-line view of one kernel straight from an ncu report that was captured with --import-source on
(no built library needed: the report carries source and SASS): the lines with the most warp-stall samples, their
executed warp instructions per particle and their shared-memory wavefronts.

    python tools/ncu_source_lines.py gpurun_out/r02f_full.ncu-rep k_deposit 134217728 [launch_index] [top] > profiles/x.md
"""
import csv
import io
import subprocess
import sys

rep, kern, npart = sys.argv[1], sys.argv[2], float(sys.argv[3])
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# the page is a sequence of (File Path, Function Name, header, rows...) blocks per source file and launch
blocks, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Function Name":
        cur["fn"] = r[1]
    elif r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and len(r) == len(cur["hdr"]) and r[0].strip().isdigit():
        cur["rows"].append(r)
# launches repeat the same set of files: split when a file path repeats
launches, seen = [], set()
for b in blocks:
    if b["file"] in seen:
        seen = set()
    if not seen:
        launches.append([])
    seen.add(b["file"])
    launches[-1].append(b)
sel = launches[min(which, len(launches) - 1)]
lines = []
for b in sel:
    ix = {h: i for i, h in enumerate(b["hdr"])}

    def f(r, k):
        try:
            return float(r[ix[k]] or 0)
        except (ValueError, KeyError):
            return 0.0
    for r in b["rows"]:
        lines.append((f(r, "# Samples"), f(r, "Instructions Executed"), f(r, "L1 Wavefronts Shared"),
                      f(r, "L1 Wavefronts Shared Ideal"), b["file"].split("/")[-1] + ":" + r[0], r[1].strip()))
tot_s = sum(l[0] for l in lines) or 1.0
tot_i = sum(l[1] for l in lines)
tot_w = sum(l[2] for l in lines)
print(f"# {sel[0].get('fn', kern)}\n")
print(f"`{rep.split('/')[-1]}`, launch {which} of this kernel: {tot_i / npart:.1f} executed warp instructions per particle, "
      f"{tot_w / npart * 32:.0f} shared-memory wavefronts per 32 particles, {tot_s:.0f} warp-stall samples.  "
      f"Lines by share of the stall samples (`tools/ncu_source_lines.py`):\n")
print("| line | stall samples | warp instr / particle | smem wavefronts / 32 particles (ideal) | source |")
print("|---|---|---|---|---|")
for s_, i_, w_, wi_, where, src in sorted(lines, reverse=True)[:top]:
    src = src.replace("|", "\\|")
    print(f"| `{where}` | {100 * s_ / tot_s:.1f} % | {i_ / npart:.2f} | {w_ / npart * 32:.1f} ({wi_ / npart * 32:.1f}) | `{src[:110]}` |")
