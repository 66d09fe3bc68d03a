"""GPU parity of the output side of the path (csrc/diag.cu) against the oracle,
which is pinned bit for bit to the reference's own XtensorPacker3D / append_moment3d / XtensorHaloMoment3D /
shape_mc<4> / shape_wt<1..4> (tests/test_oracle_vs_ref.py)."""
import numpy as np
import pytest

from nix_b200 import core
from nix_b200.synth import Problem
from oracle import nixoracle as no

from helpers import gpu_domain, oracle_domain

pytestmark = pytest.mark.gpu
PD = no.C.POINTER(no.C.c_double)


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_shape_functions_bit_exact(oracle_port, gpu_lib, order):
    rng = np.random.default_rng(order)
    n = 2000
    X = rng.uniform(-5, 5, n)
    x = X + rng.uniform(-0.5, 1.0, n)
    rdx, dt = 1.25, 0.37
    ref_mc, ref_wt = np.zeros((n, order + 1)), np.zeros((n, order + 1))
    s = np.zeros(order + 1)
    for i in range(n):
        oracle_port.nixo_shape_mc(order, x[i], X[i], rdx, s.ctypes.data_as(PD))
        ref_mc[i] = s
        oracle_port.nixo_shape_wt(order, x[i], X[i], rdx, dt, 1 / dt, s.ctypes.data_as(PD))
        ref_wt[i] = s
    assert np.array_equal(core.shape_eval(0, order, x, X, rdx).view(np.int64), ref_mc.view(np.int64))
    assert np.array_equal(core.shape_eval(1, order, x, X, rdx, dt, 1 / dt).view(np.int64), ref_wt.view(np.int64))
    assert np.allclose(ref_wt.sum(axis=1), 1.0, atol=1e-12)


def _tag_tracers(prob):
    parts = prob.particles

    def tagged(k, s):
        xu = parts(k, s)
        ids = np.ascontiguousarray(xu[:, 6]).view(np.int64).copy()
        ids[::3] = -ids[::3] - 1
        xu[:, 6] = ids.view(np.float64)
        return xu
    prob.particles = tagged


@pytest.mark.parametrize("order,cdims,dims", [(2, (2, 2, 2), (8, 8, 8)), (1, (1, 2, 2), (8, 12, 16)), (3, (2, 1, 2), (8, 8, 8))])
def test_packers_and_moments(oracle_port, gpu_lib, order, cdims, dims):
    """after two steps: pack_field and pack_moment(J) bit-identical to the reference's packers for decimate 1, 2,
    4 and 'everything' (J itself agrees to 1e-12, so its packed form is compared at that level), tracers bit-exact
    and in order, moments within 1e-12 of their maximum incl. the moment halo."""
    prob = Problem(cdims, dims, order, ppc=6, seed=33 + order, vth=(0.4, 0.1))
    _tag_tracers(prob)
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True)
    for _ in range(2):
        od.step(0.5, 1.0)
        gd.step(0.5)
    assert gd.check() == 0
    od.deposit_moment(1.0)
    gd.deposit_moment()
    for k, c in enumerate(od.chunks):
        for dec in (1, 2, 4, 64):
            assert np.array_equal(gd.pack_field(k, dec), c.pack_field(dec)), f"pack_field chunk {k} decimate {dec}"
            a, b = gd.pack_moment(k, 0, dec), c.pack_moment(0, dec)
            assert a.shape == b.shape and np.abs(a - b).max() <= 1e-12 * np.abs(c.uj).max()
            a, b = gd.pack_moment(k, 1, dec), c.pack_moment(1, dec)
            assert a.shape == b.shape and np.abs(a - b).max() <= 1e-12 * np.abs(c.um).max()
        um = gd.get_moment(k)
        assert np.abs(um - c.um).max() <= 1e-12 * np.abs(c.um).max(), f"moments chunk {k}"
        for s in range(prob.ns):
            t = gd.pack_tracer(k, s)
            r = c.pack_tracer(s)
            assert len(r) > 0 and t.shape == r.shape and np.array_equal(t.view(np.int64), r.view(np.int64))
    gd.close()


def test_pack_of_uploaded_fields_is_bit_exact(oracle_port, gpu_lib):
    """the packers alone, on arrays both sides hold bit for bit (no deposit in between): field, J and moments"""
    prob = Problem((1, 1, 2), (8, 8, 8), 2, ppc=1, seed=3)
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True)
    rng = np.random.default_rng(1)
    for k, c in enumerate(od.chunks):
        c.uj[...] = rng.uniform(-1, 1, c.uj.shape)
        gd.set_current(k, c.uj)
        for dec in (1, 2, 8):
            assert np.array_equal(gd.pack_moment(k, 0, dec), c.pack_moment(0, dec))
            assert np.array_equal(gd.pack_field(k, dec), c.pack_field(dec))
    gd.close()


def test_fp32_diagnostics(oracle_port, gpu_lib):
    prob = Problem((2, 2, 1), (8, 8, 8), 2, ppc=6, seed=9, vth=(0.3, 0.05))
    _tag_tracers(prob)
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=False, fp32=True)
    od.step(0.5, 1.0)
    gd.step(0.5)
    od.deposit_moment(1.0)
    gd.deposit_moment()
    for k, c in enumerate(od.chunks):
        assert np.abs(gd.pack_field(k, 2) - c.pack_field(2)).max() < 1e-6
        assert np.abs(gd.get_moment(k) - c.um).max() <= 1e-5 * np.abs(c.um).max()
        for s in range(prob.ns):
            t, r = gd.pack_tracer(k, s), c.pack_tracer(s)
            assert t.shape == r.shape and np.array_equal(t[:, 6].view(np.int64), r[:, 6].view(np.int64))
            assert np.abs(t[:, :6] - r[:, :6]).max() < 1e-5
    gd.close()
