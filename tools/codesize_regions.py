#!/usr/bin/env python
"""Static SASS instruction count of k_push_deposit<O,S> by source region (hot-loop code must fit the
32 KB instruction cache).  python tools/codesize_regions.py [k_push_depositILi2ELb0]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fn = sys.argv[1] if len(sys.argv) > 1 else "k_push_depositILi2ELb0"
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "nix_b200", "libnixb200.so")], cwd=d, capture_output=True)
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, "push_deposit.sm_100a.cubin")], capture_output=True,
                      text=True).stdout
lines = open(os.path.join(ROOT, "nix_b200", "csrc", "push_deposit.cu")).read().split("\n")


def find(pat):
    for i, l in enumerate(lines):
        if pat in l:
            return i + 1
    return None


marks = [(find(p), n) for p, n in [
    ("void shape_mc(", "shape"), ("double gather1(", "gather1"), ("void gather_pair(", "gather_pair"),
    ("void plane_accumulate(", "plane_accumulate"), ("void plane_reduce(", "plane_reduce"),
    ("void flush_movers(", "flush_movers"), ("__global__ void __launch_bounds__", "prologue"),
    ("auto flush_bin =", "flush_bin"), ("auto advance =", "loop ctl"), ("  while (have) {", "loop head/bin start"),
    ("=============================== push", "weights/sorted"), ("---- gather: Ex Ey", "gather call"),
    ("---- push_boris", "boris/pos"), ("---- bin of the new position", "classify"),
    ("---- 1-D deposit weights", "dep weights"), ("---- leavers: ordered", "leavers"),
    ("=============================== deposit", "deposit loop"), ("---- movers: record", "movers rec"),
    ("last iteration of a bin", "leaver slab"), ("---- flush the J tile", "epilogue")]]
marks = sorted((a, b) for a, b in marks if a)


def region(f, l):
    if f == "common.cuh":
        return "common.cuh:%d" % l
    if f != "push_deposit.cu":
        return "hdr:" + f
    r = "pre"
    for a, b in marks:
        if l >= a:
            r = b
    return r


cur, infn = None, False
cnt = collections.Counter()
for l in sass.split("\n"):
    if l.startswith(".text.") or l.startswith("//--------------------- .text"):
        infn = fn in l
    if not infn:
        continue
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l) and cur:
        cnt[region(*cur)] += 1
print(sum(cnt.values()), "instructions =", sum(cnt.values()) * 16 // 1024, "KB")
for k, v in cnt.most_common(30):
    print(f"{v:6d} {k}")
