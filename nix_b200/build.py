"""Build libnixb200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m nix_b200.build [--force]

The library is built next to this file so that it travels with the source snapshot to the GPU box.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnixb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "nixb200.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in sources() + _deps())


def build(force=False, verbose=False, defines=(), out=None):
    """Incremental: every .cu is compiled to build/<name>.o (in parallel, only when it or a header
    changed), then linked.  defines / out: experiment builds (e.g. -DNIX_PUSH_MINB=3 into
    libnixb200_x.so, selected at run time with NIXB200_LIB); the default build takes neither."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tag = "_".join(sorted(defines)).replace("=", "-") if defines else "default"
    objdir = os.path.join(HERE, "build", tag)
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(p) for p in _deps())
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + ["-D" + d for d in defines]
    procs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and os.path.getmtime(obj) > hdr_t):
            continue
        procs.append((src, subprocess.Popen([nvcc] + cflags + ["-c", "-o", obj, src], stdout=subprocess.PIPE,
                                            stderr=subprocess.PIPE, text=True)))
    report, failed = [], False
    for src, p in procs:
        so, se = p.communicate()
        report.append(so + se)
        if verbose or p.returncode != 0:
            sys.stderr.write(so + se)
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libnixb200.so")
    res = subprocess.run([nvcc, "-shared", "-o", out or LIB] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libnixb200.so")
    if out is None and report:
        with open(os.path.join(HERE, "ptxas_report.txt"), "a" if len(procs) < len(objs) else "w") as f:
            f.write("".join(report))
    return out or LIB


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    out = os.path.join(HERE, outs[0]) if outs else None
    print(build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, defines=defs, out=out))
