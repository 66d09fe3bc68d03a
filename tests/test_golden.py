"""The plain-C oracle (oracle/nix_oracle.c) against the golden vectors generated from the reference
itself (tests/golden/make_golden.py ran the reference's own templates via oracle/_ref).  Bit-exact:
both sides evaluate the same scalar expressions without FMA contraction."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import nixoracle as no

from helpers import bits, load_golden_domain

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PD = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(GOLD, "primitives.npz"))


@pytest.mark.parametrize("order", [1, 2, 3])
def test_shape_mc_golden(oracle_port, prim, order):
    x, X, s = prim[f"shape{order}_x"], prim[f"shape{order}_X"], prim[f"shape{order}_s"]
    for i in range(len(x)):
        w = np.zeros(order + 1)
        oracle_port.nixo_shape_mc(order, float(x[i]), float(X[i]), 1.0, w.ctypes.data_as(PD))
        assert np.array_equal(bits(w), bits(s[i]))


def test_push_boris_golden(oracle_port, prim):
    u, eb, ref = prim["boris_u"], prim["boris_eb"], prim["boris_out"]
    for i in range(len(u)):
        v = u[i].copy()
        oracle_port.nixo_push_boris(v.ctypes.data_as(PD), np.ascontiguousarray(eb[i]).ctypes.data_as(PD), 1.0)
        assert np.array_equal(bits(v), bits(ref[i]))


@pytest.mark.parametrize("order", [1, 2, 3])
def test_deposit3d_golden(oracle_port, prim, order):
    ss, ref = prim[f"dep{order}_ss"], prim[f"dep{order}_cur"]
    for i in range(len(ss)):
        s_in = np.ascontiguousarray(ss[i]).copy()
        cur = np.zeros_like(ref[i])
        oracle_port.nixo_deposit3d(order, 2.0, 2.0, 2.0, -1.0, s_in.ctypes.data_as(PD), cur.ctypes.data_as(PD))
        assert np.abs(ref[i]).max() > 0
        assert np.array_equal(bits(cur), bits(ref[i]))


@pytest.mark.parametrize("order", [1, 2, 3])
def test_full_steps_golden(oracle_port, order):
    g = np.load(os.path.join(GOLD, f"steps_order{order}.npz"))
    dom, ns = load_golden_domain(oracle_port, g, order)
    for k, c in enumerate(dom.chunks):
        for s in range(ns):
            assert np.array_equal(bits(c.particles(s)), bits(g[f"sorted_xu_{k}_{s}"]))
            assert np.array_equal(c.pindex(s), g[f"sorted_pindex_{k}_{s}"])
    for _ in range(int(g["nstep"])):
        dom.step(float(g["delt"]), float(g["cc"]))
    for k, c in enumerate(dom.chunks):
        assert np.array_equal(bits(c.uf), bits(g[f"out_uf_{k}"]))
        assert np.array_equal(bits(c.uj), bits(g[f"out_uj_{k}"]))
        for s in range(ns):
            assert np.array_equal(bits(c.particles(s)), bits(g[f"out_xu_{k}_{s}"]))
            assert np.array_equal(c.pindex(s), g[f"out_pindex_{k}_{s}"])


@pytest.mark.parametrize("name", ["vay", "hc"])
def test_other_pushers_golden(oracle_port, prim, name):
    """push_vay / push_higuera_cary (primitives.hpp:193-253) against values written by the reference."""
    fn = oracle_port.nixo_push_vay if name == "vay" else oracle_port.nixo_push_higuera_cary
    u, eb, ref = prim["boris_u"], prim["boris_eb"], prim[f"{name}_out"]
    for i in range(len(u)):
        v = u[i].copy()
        fn(v.ctypes.data_as(PD), np.ascontiguousarray(eb[i]).ctypes.data_as(PD), 1.0)
        assert np.array_equal(bits(v), bits(ref[i]))
    assert not np.array_equal(bits(ref), bits(prim["boris_out"]))


@pytest.mark.parametrize("name", ["vay", "hc"])
def test_full_steps_other_pushers_golden(oracle_port, name):
    g = np.load(os.path.join(GOLD, f"steps_order2_{name}.npz"))
    dom, ns = load_golden_domain(oracle_port, g, 2)
    oracle_port.nixo_set_pusher(int(g["pusher"]))
    try:
        for _ in range(int(g["nstep"])):
            dom.step(float(g["delt"]), float(g["cc"]))
    finally:
        oracle_port.nixo_set_pusher(0)
    for k, c in enumerate(dom.chunks):
        assert np.array_equal(bits(c.uj), bits(g[f"out_uj_{k}"]))
        for s in range(ns):
            assert np.array_equal(bits(c.particles(s)), bits(g[f"out_xu_{k}_{s}"]))
            assert np.array_equal(c.pindex(s), g[f"out_pindex_{k}_{s}"])
