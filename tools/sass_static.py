#!/usr/bin/env python
"""Static view of every kernel in libnixb200.so, from the built library alone (no GPU needed):
registers / shared memory / spills (cuobjdump -res-usage) and how many SASS instructions of the classes that
matter on this path each kernel contains (cuobjdump -sass): TMA loads (UTMALDG), asynchronous copies
(LDGSTS), mbarrier waits (SYNCS), fp64 arithmetic (DFMA / DMUL / DADD), fp32 arithmetic, shared-memory
loads / stores / atomics (ATOMS: the compare-and-swap loops of the fp64 shared-memory adds show up as
ATOMS.CAST.SPIN), global atomics and reductions (ATOMG / RED), warp shuffles / votes / matches.

    python tools/sass_static.py [lib.so] > profiles/rNN_sass_static.md

Static counts are NOT execution counts (loops, predication); executed counts come from ncu
(tools/ncu_hot.py).  What this table proves is which hardware paths a kernel uses at all.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nix_b200", "libnixb200.so")

CLASSES = [
    ("UTMALDG", r"^UTMALDG"), ("LDGSTS", r"^LDGSTS"), ("SYNCS", r"^SYNCS"),
    ("DFMA", r"^DFMA"), ("DMUL", r"^DMUL"), ("DADD", r"^DADD"), ("D-other", r"^(DSETP|MUFU\.RCP64H|MUFU\.RSQ64H|F2F\.F64|I2F\.F64|F2I\.\S*F64)"),
    ("FFMA", r"^FFMA"), ("FMUL/FADD", r"^(FMUL|FADD)"),
    ("LDS", r"^LDS"), ("STS", r"^STS"), ("ATOMS", r"^ATOMS"), ("ATOMS.CAS", r"^ATOMS\.CAST?"),
    ("ATOMG", r"^ATOMG"), ("RED", r"^RED"), ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDG/STG.128", r"^(LDG|STG)\.\S*128"),
    ("SHFL", r"^SHFL"), ("VOTE/MATCH", r"^(VOTE|MATCH|REDUX)"), ("BAR", r"^BAR"), ("LDL/STL", r"^(LDL|STL)"),
]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    res = {}
    for n, d in zip(names, out):
        d = re.sub(r"nixb200::(\(anonymous namespace\)|<unnamed>)::", "", d)
        d = re.sub(r"nixb200::", "", d)
        d = re.sub(r"\(.*$", "", d)  # drop the argument list
        d = re.sub(r"^void ", "", d)
        res[n] = d
    return res


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout.split("\n")
    usage, cur = {}, None
    for l in res:
        m = re.match(r"\s*Function (\S+):", l)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in l:
            d = dict(re.findall(r"(\w+):(\d+)", l))
            usage[cur] = d
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
    counts, total, cur = collections.defaultdict(collections.Counter), collections.Counter(), None
    for l in sass:
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]+)", l)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            for name, pat in CLASSES:
                if re.match(pat, op):
                    counts[cur][name] += 1
    names = sorted(total, key=lambda n: -total[n])
    dm = demangle(names)
    cols = [c for c, _ in CLASSES]
    print(f"# Static SASS view of `{os.path.relpath(lib, ROOT)}` (sm_100a), `tools/sass_static.py`\n")
    print("Registers / shared memory / local memory per thread from `cuobjdump -res-usage`; instruction classes counted in "
          "`cuobjdump -sass` (static occurrences, not executions).\n")
    print("| kernel | regs | smem static B | local B | SASS instr | " + " | ".join(cols) + " |")
    print("|---|---|---|---|---|" + "---|" * len(cols))
    for n in names:
        u = usage.get(n, {})
        row = [dm[n], u.get("REG", "?"), u.get("SHARED", "?"), u.get("LOCAL", "?"), str(total[n])]
        row += [str(counts[n][c]) if counts[n][c] else "" for c in cols]
        print("| `" + row[0] + "` | " + " | ".join(row[1:]) + " |")
    # the two statements DESIGN.md makes about the instruction set in use
    f32 = [n for n in names if re.search(r"float|, f>|Ef|If", dm[n]) and ("k_push" in dm[n] or "k_deposit" in dm[n])]
    bad = [dm[n] for n in f32 if counts[n]["DFMA"] + counts[n]["DMUL"] + counts[n]["DADD"] > 0]
    print(f"\nfp32 instantiations of k_push / k_deposit containing an fp64 arithmetic instruction: {bad if bad else 'none'} "
          f"(of {len(f32)})")
    tma = [dm[n] for n in names if counts[n]["UTMALDG"]]
    print(f"\nkernels issuing TMA tensor loads (UTMALDG): {len(tma)}: " + ", ".join(f"`{t}`" for t in tma))


if __name__ == "__main__":
    main()
