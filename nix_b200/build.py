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


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "nixb200.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: experiment builds (e.g. -DNIX_PUSH_MINB=3 into libnixb200_x.so, selected at run
    time with NIXB200_LIB); the default build takes neither."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out or LIB] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libnixb200.so")
    if out is None:
        with open(os.path.join(HERE, "ptxas_report.txt"), "w") as f:
            f.write(res.stderr)
    return out or LIB


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    out = os.path.join(HERE, outs[0]) if outs else None
    print(build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, defines=defs, out=out))
