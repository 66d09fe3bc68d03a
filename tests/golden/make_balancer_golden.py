"""Generate tests/golden/balancer.npz from the reference's OWN nix::Balancer (balancer.cpp), through
host/_build/demo (which compiles balancer.cpp where it lies under /root/reference):

    make -C host && python tests/golden/make_balancer_golden.py
"""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DEMO = os.path.join(ROOT, "host", "_build", "demo")


def run(nrank, mode, load, boundary=None):
    cmd = [DEMO, "assign", str(nrank), mode] + [repr(float(v)) for v in load]
    if boundary is not None:
        cmd += ["--"] + [str(int(b)) for b in boundary]
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    return np.array([int(v) for v in out.split()], dtype=np.int32)


rng = np.random.default_rng(0)
out = {}
case = 0
for nchunk, nrank in [(16, 2), (16, 4), (64, 8), (512, 8), (100, 7), (12, 12), (40, 3)]:
    for kind in ("uniform", "random", "sine", "spike"):
        if kind == "uniform":
            load = np.ones(nchunk)
        elif kind == "random":
            load = rng.uniform(0.1, 2.0, nchunk)
        elif kind == "sine":
            load = 1.0 + 0.8 * np.sin(2 * np.pi * (np.arange(nchunk) + 0.5) / nchunk)
        else:
            load = np.ones(nchunk)
            load[nchunk // 3] = 0.4 * nchunk
        b0 = run(nrank, "initial", load)
        uni = np.array([(nchunk * r) // nrank for r in range(nrank + 1)], dtype=np.int32)
        b1 = run(nrank, "step", load, uni)
        b2 = run(nrank, "step", load, b1)
        out[f"load_{case}"] = load
        out[f"nrank_{case}"] = np.int32(nrank)
        out[f"initial_{case}"] = b0
        out[f"uniform_{case}"] = uni
        out[f"step1_{case}"] = b1
        out[f"step2_{case}"] = b2
        case += 1
out["ncase"] = np.int32(case)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "balancer.npz"), **out)
print("wrote", case, "cases")
