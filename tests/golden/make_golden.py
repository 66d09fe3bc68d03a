#!/usr/bin/env python
"""Generate the golden vectors of tests/golden/ FROM THE REFERENCE ITSELF.

    python tests/golden/make_golden.py        # needs /root/reference (builds oracle/_ref)

Every output below is produced by oracle/_ref/libnixref.so, i.e. by the reference's OWN templates
and classes (primitives.hpp, interp.hpp, esirkepov.hpp, xtensor_particle.hpp, xtensor_halo3d.hpp,
chunk.cpp) compiled from /root/reference with -ffp-contract=off and driven by
oracle/ref/ref_driver.cpp.  The inputs are stored next to the outputs so that the fixtures do not
depend on any random-number generator at test time.  The fixtures pin
  * the plain-C restatement oracle/nix_oracle.c    (tests/test_golden.py, CPU), and
  * the CUDA path through the C ABI                 (tests/test_gpu_golden.py, GPU)
on machines where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from nix_b200.synth import Problem  # noqa: E402
from oracle import nixoracle as no  # noqa: E402

NSTEP = 3
DELT, CC = 0.5, 1.0


def primitives(lib):
    rng = np.random.default_rng(20261017)
    out = {}
    # shape_mc<1..3>  (primitives.hpp:257-298,519-532)
    for order in (1, 2, 3):
        x = rng.uniform(2.0, 9.0, 64)
        rdx = 1.0
        if order % 2 == 1:
            X = np.floor(x)
        else:
            X = np.floor(x + 0.5)
        s = np.zeros((64, order + 1))
        for i in range(64):
            buf = np.zeros(order + 1)
            lib.nixo_shape_mc(order, float(x[i]), float(X[i]), rdx, buf.ctypes.data_as(no.C.POINTER(no.C.c_double)))
            s[i] = buf
        out[f"shape{order}_x"], out[f"shape{order}_X"], out[f"shape{order}_s"] = x, X, s
    # push_boris (primitives.hpp:165-189): the reference's scalar test section is empty, so these
    # values pin it
    u = rng.normal(0, 1.0, (32, 3))
    eb = rng.normal(0, 0.3, (32, 6))
    uo = u.copy()
    for i in range(32):
        lib.nixo_push_boris(uo[i].ctypes.data_as(no.C.POINTER(no.C.c_double)),
                            eb[i].ctypes.data_as(no.C.POINTER(no.C.c_double)), CC)
    out["boris_u"], out["boris_eb"], out["boris_out"] = u, eb, uo
    # push_vay / push_higuera_cary (primitives.hpp:193-253) on the same inputs
    for name, fn in (("vay", lib.nixo_push_vay), ("hc", lib.nixo_push_higuera_cary)):
        uo = u.copy()
        for i in range(32):
            fn(uo[i].ctypes.data_as(no.C.POINTER(no.C.c_double)), eb[i].ctypes.data_as(no.C.POINTER(no.C.c_double)), CC)
        out[f"{name}_out"] = uo
    # deposit3d<1..3> (esirkepov.hpp:326-340) on random valid weight sets
    for order in (1, 2, 3):
        ns = order + 3
        ss = np.zeros((8, 2, 3, ns))
        cur = np.zeros((8, ns, ns, ns, 4))
        for i in range(8):
            for d in range(3):
                x0 = rng.uniform(3.0, 4.0)
                x1 = x0 + rng.uniform(-0.45, 0.45)
                for t, xx in enumerate((x0, x1)):
                    if order % 2 == 1:
                        i0 = int(np.floor(x0))
                        i1 = int(np.floor(xx))
                    else:
                        i0 = int(np.floor(x0 + 0.5))
                        i1 = int(np.floor(xx + 0.5))
                    w = np.zeros(order + 1)
                    lib.nixo_shape_mc(order, float(xx), float(i1), 1.0, w.ctypes.data_as(no.C.POINTER(no.C.c_double)))
                    ss[i, t, d, 1 + (i1 - i0):1 + (i1 - i0) + order + 1] = w
            s_in = ss[i].copy()
            lib.nixo_deposit3d(order, 2.0, 2.0, 2.0, -1.0, s_in.ctypes.data_as(no.C.POINTER(no.C.c_double)),
                               cur[i].ctypes.data_as(no.C.POINTER(no.C.c_double)))
        out[f"dep{order}_ss"], out[f"dep{order}_cur"] = ss, cur
    return out


def full_steps(lib, order, pusher=0):
    """pusher: 0 push_boris, 1 push_vay, 2 push_higuera_cary (the composed step of oracle/ref/ref_driver.cpp)"""
    lib.nixo_set_pusher(pusher)
    prob = Problem((2, 1, 2), (4, 4, 4), order, ppc=2, seed=900 + order + 10 * pusher, vth=(0.4, 0.1))
    dom = no.Domain(lib, prob.cdims, prob.dims, prob.nb, prob.order, prob.ns, prob.q, prob.m, prob.coord,
                    prob.ncell() * prob.ppc)
    out = {"cdims": np.array(prob.cdims), "dims": np.array(prob.dims), "nb": np.array(prob.nb),
           "coord": prob.coord, "q": prob.q, "m": prob.m, "nstep": np.array(NSTEP),
           "delt": np.array(DELT), "cc": np.array(CC)}
    for k, c in enumerate(dom.chunks):
        uf = prob.field(k)
        out[f"in_uf_{k}"] = uf
        c.uf[...] = uf
        for s in range(prob.ns):
            xu = prob.particles(k, s)
            out[f"in_xu_{k}_{s}"] = xu
            c.set_particles(s, xu)
    dom.exchange(no.MODE_FIELD)
    dom.sort_only()
    for k, c in enumerate(dom.chunks):
        for s in range(prob.ns):
            out[f"sorted_xu_{k}_{s}"] = c.particles(s)
            out[f"sorted_pindex_{k}_{s}"] = c.pindex(s).copy()
    for _ in range(NSTEP):
        dom.step(DELT, CC, False)
    lib.nixo_set_pusher(0)
    out["pusher"] = np.array(pusher)
    for k, c in enumerate(dom.chunks):
        out[f"out_uf_{k}"] = c.uf.copy()
        out[f"out_uj_{k}"] = c.uj.copy()
        for s in range(prob.ns):
            out[f"out_xu_{k}_{s}"] = c.particles(s)
            out[f"out_pindex_{k}_{s}"] = c.pindex(s).copy()
    return out


def main():
    if not os.path.isdir("/root/reference"):
        raise SystemExit("make_golden.py needs the reference sources at /root/reference")
    no.build(("ref",))
    lib = no.load("ref")
    assert lib.nixo_impl_name().decode() == "reference"
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **primitives(lib))
    for order in (1, 2, 3):
        np.savez_compressed(os.path.join(HERE, f"steps_order{order}.npz"), **full_steps(lib, order))
    for pusher, name in ((1, "vay"), (2, "hc")):
        np.savez_compressed(os.path.join(HERE, f"steps_order2_{name}.npz"), **full_steps(lib, 2, pusher))
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
