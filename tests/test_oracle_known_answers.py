"""The reference's own known-answer and invariant tests, restated against the oracle
(SURVEY.md section 4 / 8c).  Each test cites the reference test it follows."""
import ctypes as C

import numpy as np
import pytest

from oracle import nixoracle as no

PD = C.POINTER(C.c_double)


def W(order, t):
    """Closed-form B-spline (unittest/test_primitives.cpp:271-530 compares shape_mc with these)."""
    a = abs(t)
    if order == 1:
        return max(0.0, 1 - a)
    if order == 2:
        if a < 0.5:
            return 0.75 - a * a
        if a < 1.5:
            return 0.5 * (1.5 - a) ** 2
        return 0.0
    if order == 3:
        if a < 1:
            return 2 / 3.0 - a * a + 0.5 * a ** 3
        if a < 2:
            return (2 - a) ** 3 / 6.0
        return 0.0
    raise ValueError(order)


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("dx", [0.5, 1.0, 1.5])
def test_shape_mc_equals_analytic(oracle_port, order, dx):
    # test_primitives.cpp:271-530: 100 points, abs 1e-14
    rng = np.random.default_rng(order)
    rdx = 1 / dx
    for x in rng.uniform(0, 10 * dx, 100):
        if order % 2:
            ix = int(np.floor(x * rdx))            # lower node   (primitives.hpp:497-511)
        else:
            ix = int(np.floor(x * rdx + 0.5))      # nearest node
        X = ix * dx
        s = np.zeros(order + 1)
        oracle_port.nixo_shape_mc(order, float(x), float(X), rdx, s.ctypes.data_as(PD))
        first = ix - (order // 2)                  # node of s[0]
        for j in range(order + 1):
            assert abs(s[j] - W(order, (x - (first + j) * dx) * rdx)) < 1e-14
        assert abs(s.sum() - 1) < 1e-14


def _tp(n, x):
    """(x)_+^n in extended precision"""
    x = np.longdouble(x)
    return x ** n if x > 0 else np.longdouble(0)


def bspline(n, t):
    """Centred cardinal B-spline of degree n (support n + 1), textbook truncated-power form
    B_n(t) = 1/n! sum_k (-1)^k C(n+1, k) (t + (n+1)/2 - k)_+^n, evaluated in extended precision.  Independent of
    the piecewise polynomials the reference's tests spell out (test_primitives.cpp:271-530) and equal to them."""
    from math import comb, factorial
    return float(sum(np.longdouble((-1) ** k * comb(n + 1, k)) * _tp(n, np.longdouble(t) + np.longdouble(n + 1) / 2 - k)
                     for k in range(n + 2)) / np.longdouble(factorial(n)))


def bspline_cumulative(m, t):
    """integral of B_m from -inf to t"""
    from math import comb, factorial
    return sum(np.longdouble((-1) ** k * comb(m + 1, k)) * _tp(m + 1, np.longdouble(t) + np.longdouble(m + 1) / 2 - k)
               for k in range(m + 2)) / np.longdouble(factorial(m + 1))


def test_truncated_power_form_equals_the_piecewise_b_splines():
    for order in (1, 2, 3):
        for t in np.linspace(-2.6, 2.6, 521):
            assert abs(bspline(order, t) - W(order, t)) < 1e-15


@pytest.mark.parametrize("dx", [0.5, 1.0, 1.5])
def test_shape_mc4_equals_analytic(oracle_port, dx):
    # test_primitives.cpp:458-530 ("Fourth-order shape function"): 100 points, abs 1e-14
    rdx = 1 / dx
    for x in np.linspace(0, 3 * dx, 100):
        ix = int(np.floor(x * rdx + 0.5))  # even order: nearest node (primitives.hpp:497-511)
        s = np.zeros(5)
        oracle_port.nixo_shape_mc(4, float(x), float(ix * dx), rdx, s.ctypes.data_as(PD))
        for j in range(5):
            assert abs(s[j] - bspline(4, (x - (ix - 2 + j) * dx) * rdx)) < 1e-14
        assert abs(s.sum() - 1) < 1e-14


@pytest.mark.parametrize("order", [1, 2, 3, 4])
@pytest.mark.parametrize("dt", [0.1, 0.2, 0.3, 0.4, 0.5])
@pytest.mark.parametrize("dx", [0.5, 1.0, 1.5])
def test_shape_wt_equals_analytic(oracle_port, order, dt, dx):
    """test_primitives.cpp:532-822 ("... shape function for WT scheme", delt = 0.1 .. 0.5, 100 points, abs 1e-14).
    The reference's tests spell the WT weights out as piecewise polynomials per order; all of them are ONE
    statement: the order-n WT weight is the B-spline of degree n-1 averaged over the path of half-length dt
    (in cells), W_n(t; dt) = [C_{n-1}(t + dt) - C_{n-1}(t - dt)] / (2 dt) with C the cumulative B-spline -- e.g.
    order 1: (1 + 2 dt - 2|t|) / (4 dt) on 1/2 - dt < |t| <= 1/2 + dt, 1 inside, 0 outside (:541-551)."""
    rdx = 1 / dx
    if order == 1:  # the reference's own closed form for order 1, as a check of the statement above
        for t in np.linspace(-1.2, 1.2, 49):
            a = abs(t)
            w1 = (1 + 2 * dt - 2 * a) / (4 * dt) if 0.5 - dt < a <= 0.5 + dt else (1.0 if a <= 0.5 - dt else 0.0)
            w = float((bspline_cumulative(0, t + dt) - bspline_cumulative(0, t - dt)) / (2 * np.longdouble(dt)))
            assert abs(w - w1) < 1e-15
    for x in np.linspace(0, 3 * dx, 100):
        ix = int(np.floor(x * rdx)) if order % 2 else int(np.floor(x * rdx + 0.5))
        s = np.zeros(order + 1)
        oracle_port.nixo_shape_wt(order, float(x), float(ix * dx), rdx, dt, 1 / dt, s.ctypes.data_as(PD))
        first = ix - order // 2
        for j in range(order + 1):
            t = (np.longdouble(x) - (first + j) * np.longdouble(dx)) * np.longdouble(rdx)
            w = float((bspline_cumulative(order - 1, t + dt) - bspline_cumulative(order - 1, t - dt)) / (2 * np.longdouble(dt)))
            assert abs(s[j] - w) < 1e-14, (order, dt, dx, x, j)
        assert abs(s.sum() - 1) < 1e-14


@pytest.mark.parametrize("xmin,dx", [(-1.0, 0.5), (0.0, 1.0), (-1.0, 1.5), (0.0, 0.5)])
def test_digitize(oracle_port, xmin, dx):
    # test_primitives.cpp:73-117
    rng = np.random.default_rng(3)
    for i in rng.integers(0, 50, 100):
        x = xmin + (i + rng.uniform(0.01, 0.99)) * dx
        assert oracle_port.nixo_digitize(float(x), xmin, 1 / dx) == i


def test_lorentz_and_boris_energy(oracle_port):
    # push_boris with E = 0 is a pure rotation: |u| is conserved (primitives.hpp:165-189)
    rng = np.random.default_rng(5)
    for _ in range(50):
        u = rng.normal(0, 1, 3)
        eb = np.concatenate([np.zeros(3), rng.normal(0, 0.4, 3)])
        v = u.copy()
        oracle_port.nixo_push_boris(v.ctypes.data_as(PD), eb.ctypes.data_as(PD), 1.0)
        assert abs(np.dot(v, v) - np.dot(u, u)) < 1e-13 * np.dot(u, u)
        g = oracle_port.nixo_lorentz_factor(float(u[0]), float(u[1]), float(u[2]), 1.0)
        assert abs(g - np.sqrt(1 + np.dot(u, u))) < 1e-14 * g


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_interp3d_equals_naive_sum(oracle_port, order):
    # test_interp.cpp:726-784: factorised result == naive triple sum, rel 1e-14
    rng = np.random.default_rng(order + 10)
    M = 12
    eb = np.ascontiguousarray(rng.uniform(-1, 1, (M, M, M, 6)))
    nw = order + 2
    for _ in range(20):
        w = [np.ascontiguousarray(rng.uniform(0, 1, nw)) for _ in range(3)]
        i0 = rng.integers(0, M - nw, 3)
        ik = int(rng.integers(0, 6))
        dt = 0.37
        got = oracle_port.nixo_interp3d(order, eb.ctypes.data_as(PD), M, M, int(i0[0]), int(i0[1]), int(i0[2]), ik,
                                        w[0].ctypes.data_as(PD), w[1].ctypes.data_as(PD), w[2].ctypes.data_as(PD), dt)
        sub = eb[i0[0]:i0[0] + nw, i0[1]:i0[1] + nw, i0[2]:i0[2] + nw, ik]
        ref = np.einsum("zyx,z,y,x->", sub, w[0], w[1], w[2]) * dt
        assert abs(got - ref) < 1e-13 * max(1.0, abs(ref))


def _weights(lib, order, x0, x1):
    ns = order + 3
    ss = np.zeros((2, 3, ns))
    for d in range(3):
        for t, xx in enumerate((x0[d], x1[d])):
            f = (lambda v: int(np.floor(v))) if order % 2 else (lambda v: int(np.floor(v + 0.5)))
            i0, i1 = f(x0[d]), f(xx)
            w = np.zeros(order + 1)
            lib.nixo_shape_mc(order, float(xx), float(i1), 1.0, w.ctypes.data_as(PD))
            ss[t, d, 1 + (i1 - i0):2 + (i1 - i0) + order] = w
    return ss


@pytest.mark.parametrize("order", [1, 2, 3, 4])
@pytest.mark.parametrize("dt,dh", [(0.5, 1.0), (1.0, 1.0), (0.5, 2.0)])
def test_deposit3d_continuity(oracle_port, order, dt, dh):
    """test_esirkepov.cpp:993-1106: (1) sum of weights = 1, (2) sum rho = q per particle,
    (3) discrete continuity  rho_new - rho_old + dt/dh (dJx + dJy + dJz) = 0  to round-off."""
    rng = np.random.default_rng(order * 7 + 1)
    ns = order + 3
    q = -1.3
    cur = np.zeros((ns + 1, ns + 1, ns + 1, 4))
    rho_old = np.zeros((ns + 1, ns + 1, ns + 1))
    npart = 64
    for _ in range(npart):
        x0 = rng.uniform(3.0, 4.0, 3)                  # (z, y, x) in cell units
        x1 = x0 + rng.uniform(-0.45, 0.45, 3)
        ss = _weights(oracle_port, order, x0, x1)
        assert np.allclose(ss.sum(axis=2), 1.0, atol=1e-13)
        rho_old[:ns, :ns, :ns] += q * np.einsum("z,y,x->zyx", ss[0, 0], ss[0, 1], ss[0, 2])
        c = np.zeros((ns, ns, ns, 4))
        # weights are passed as ss[t][dir] with dir order (x, y, z) in the reference's deposit3d
        s_in = np.ascontiguousarray(ss[:, ::-1, :]).copy()
        oracle_port.nixo_deposit3d(order, dh / dt, dh / dt, dh / dt, q, s_in.ctypes.data_as(PD), c.ctypes.data_as(PD))
        cur[:ns, :ns, :ns] += c
    rho_new = cur[..., 0]
    assert abs(rho_new.sum() - q * npart) < 1e-12 * abs(q * npart)
    jx, jy, jz = cur[..., 1], cur[..., 2], cur[..., 3]
    # forward differences: J[j+1] is the flux between nodes j and j+1 (test_esirkepov.cpp:1018-1022)
    div = np.zeros_like(rho_new)
    div[:, :, :-1] += jx[:, :, 1:] - jx[:, :, :-1]
    div[:, :-1, :] += jy[:, 1:, :] - jy[:, :-1, :]
    div[:-1, :, :] += jz[1:, :, :] - jz[:-1, :, :]
    res = (rho_new - rho_old) + dt / dh * div
    assert np.abs(res).sum() < 1e-13 * np.abs(rho_new).sum()


@pytest.mark.parametrize("nb", [1, 2, 3])
def test_halo_field_plus_x(oracle_port, nb):
    # test_xtensor_halo3d.cpp:51-107: ghost == source interior, exact
    c = no.Chunk(oracle_port, (4, 4, 4), nb, 1)
    for s in range(27):
        c.set_nb_valid(s // 9, (s // 3) % 3, s % 3, False)
    c.set_nb_valid(1, 1, 2, True)
    uf = c.uf
    uf[...] = -1.0
    Lb, Ub = nb, nb + 3
    iz, iy, ix, k = np.meshgrid(np.arange(Lb, Ub + 1), np.arange(Lb, Ub + 1), np.arange(Lb, Ub + 1), np.arange(6),
                                indexing="ij")
    uf[Lb:Ub + 1, Lb:Ub + 1, Lb:Ub + 1, :] = iz * 100000 + iy * 1000 + ix * 10 + k
    c.halo_pack(no.MODE_FIELD)
    size, addr = c.bufsize(no.MODE_FIELD), c.bufaddr(no.MODE_FIELD)
    e = 9 * 1 + 3 * 1 + 2
    assert size[e] == 48 * 4 * 4 * nb                  # chunk.cpp:268-272
    c.recvbuf(no.MODE_FIELD)[addr[e]:addr[e] + size[e]] = c.sendbuf(no.MODE_FIELD)[addr[e]:addr[e] + size[e]]
    c.halo_unpack(no.MODE_FIELD)
    for layer in range(nb):
        src, dst = Ub - nb + 1 + layer, Ub + 1 + layer
        exp = (iz[:, :, 0, :] * 100000 + iy[:, :, 0, :] * 1000 + src * 10 + k[:, :, 0, :]).astype(float)
        assert np.array_equal(uf[Lb:Ub + 1, Lb:Ub + 1, dst, :], exp)
    # nothing else was touched
    assert (uf[:, :, :Lb, :] == -1).all() and (uf[:Lb] == -1).all()


@pytest.mark.parametrize("nb", [1, 2, 3])
def test_halo_current_plus_x_adds(oracle_port, nb):
    # test_xtensor_halo3d.cpp:109-162
    c = no.Chunk(oracle_port, (4, 4, 4), nb, 1)
    for s in range(27):
        c.set_nb_valid(s // 9, (s // 3) % 3, s % 3, False)
    c.set_nb_valid(1, 1, 2, True)
    uj = c.uj
    uj[...] = 0.0
    Lb, Ub = nb, nb + 3
    uj[Lb:Ub + 1, Lb:Ub + 1, Ub + 1:Ub + nb + 1, :] = 2.5
    c.halo_pack(no.MODE_CURRENT)
    size, addr = c.bufsize(no.MODE_CURRENT), c.bufaddr(no.MODE_CURRENT)
    e = 9 + 3 + 2
    c.recvbuf(no.MODE_CURRENT)[addr[e]:addr[e] + size[e]] = c.sendbuf(no.MODE_CURRENT)[addr[e]:addr[e] + size[e]]
    c.halo_unpack(no.MODE_CURRENT)
    assert np.array_equal(uj[Lb:Ub + 1, Lb:Ub + 1, Ub - nb + 1:Ub + 1, :], np.full((4, 4, nb, 4), 2.5))


@pytest.mark.parametrize("nb", [1, 2, 3])
def test_halo_particle_plus_x_wraps(oracle_port, nb):
    # test_xtensor_halo3d.cpp:164-215: the out-of-range particle is wrapped and kept, Np == 2
    c = no.Chunk(oracle_port, (4, 4, 4), nb, 1, ns=1, np_required=[4])
    for s in range(27):
        c.set_nb_valid(s // 9, (s // 3) % 3, s % 3, False)
    c.set_nb_valid(1, 1, 2, True)
    xu = np.zeros((2, 7))
    xu[0, :3] = 0.5
    xu[1, :3] = (4.0 + 0.1, 0.5, 0.5)
    c.set_particles(0, xu)
    c.count(0, 0, 1, True)
    c.halo_pack(no.MODE_PARTICLE)
    size = c.bufsize(no.MODE_PARTICLE)
    e = 9 + 3 + 2
    assert size[e] == 4 + 56                           # xtensor_halo3d.hpp:258-259,320
    send = c.sendbuf(no.MODE_PARTICLE).copy()
    saddr = c.bufaddr(no.MODE_PARTICLE).copy()
    c.set_recv_sizes(no.MODE_PARTICLE, size)
    raddr = c.bufaddr(no.MODE_PARTICLE)
    c.recvbuf(no.MODE_PARTICLE)[raddr[e]:raddr[e] + size[e]] = send[saddr[e]:saddr[e] + size[e]]
    c.halo_unpack(no.MODE_PARTICLE)
    assert c.np(0) == 2
    p = c.particles(0)
    assert ((p[:, 0] >= 0) & (p[:, 0] < 4.0)).all()


def test_sort_is_stable_by_cell_and_lane(oracle_port):
    """the net effect of count+sort is a stable sort by cell*8 + ip%8 with
    out-of-bounds particles dropped (xtensor_particle.hpp:260-357) -- the reference's own sort
    tests are vacuous (Np == 0), so this property is pinned here and in test_oracle_vs_ref.py."""
    rng = np.random.default_rng(8)
    for order in (1, 2, 3):
        nb = 3 if order == 3 else 2
        c = no.Chunk(oracle_port, (5, 6, 7), nb, order, ns=1, np_required=[3000])
        n = 3000
        xu = np.zeros((n, 7))
        xu[:, 0] = rng.uniform(-0.5, 7.5, n)
        xu[:, 1] = rng.uniform(0, 6, n)
        xu[:, 2] = rng.uniform(0, 5, n)
        xu[:, 6] = np.arange(n)
        c.set_particles(0, xu)
        c.count(0, 0, n - 1, True)
        gi = c.gindex(0)[:n].copy()
        ng = c.ng(0)
        c.sort(0)
        inb = gi < ng
        assert (inb == ((xu[:, 0] >= 0) & (xu[:, 0] < 7))).all()
        key = gi.astype(np.int64) * 8 + (np.arange(n) % 8)
        perm = np.argsort(key[inb], kind="stable")
        exp = xu[inb][perm]
        assert c.np(0) == inb.sum()
        assert np.array_equal(c.particles(0), exp)
        assert c.pindex(0)[ng] == inb.sum()


# ---- XtensorPacker3D: the reference's own known answers (unittest/test_xtensor_packer3d.cpp) --------------
def _chunk_with_pattern(lib, n, nb):
    from oracle import nixoracle as no
    c = no.Chunk(lib, (n, n, n), nb, 2, ns=1, np_required=[64])
    iz, iy, ix, k = np.meshgrid(*[np.arange(m) for m in c.M], np.arange(6), indexing="ij")
    c.uf[...] = iz * 1000.0 + iy * 100.0 + ix * 10.0 + k   # test_xtensor_packer3d.cpp:191-199
    c.uj[...] = (iz * 100.0 + iy * 10.0 + ix)[..., :4]       # :235-242
    return c


def test_pack_field_colocates(oracle_port):
    """test_xtensor_packer3d.cpp:188-230: every packed value is the average of the staggered neighbours"""
    n, nb = 4, 2
    c = _chunk_with_pattern(oracle_port, n, nb)
    out = c.pack_field(1).reshape(n, n, n, 6)
    x = c.uf
    s = slice(nb, nb + n)
    p = slice(nb + 1, nb + n + 1)
    assert np.allclose(out[..., 0], 0.5 * (x[s, s, s, 0] + x[s, s, p, 0]))
    assert np.allclose(out[..., 1], 0.5 * (x[s, s, s, 1] + x[s, p, s, 1]))
    assert np.allclose(out[..., 2], 0.5 * (x[s, s, s, 2] + x[p, s, s, 2]))
    assert np.allclose(out[..., 3], 0.25 * (x[s, s, s, 3] + x[p, p, s, 3] + x[s, p, s, 3] + x[p, s, s, 3]))
    assert np.allclose(out[..., 4], 0.25 * (x[s, s, s, 4] + x[p, s, p, 4] + x[p, s, s, 4] + x[s, s, p, 4]))
    assert np.allclose(out[..., 5], 0.25 * (x[s, s, s, 5] + x[s, p, p, 5] + x[s, s, p, 5] + x[s, p, s, 5]))


def test_pack_moment_decimates_by_averaging_blocks(oracle_port):
    """test_xtensor_packer3d.cpp:232-275: decimate = 2 on a 4^3 interior -> 2^3 block means; decimate >= size -> 1"""
    n, nb = 4, 2
    c = _chunk_with_pattern(oracle_port, n, nb)
    out = c.pack_moment(0, 2).reshape(2, 2, 2, 4)
    inner = c.uj[nb:nb + n, nb:nb + n, nb:nb + n]
    want = inner.reshape(2, 2, 2, 2, 2, 2, 4).mean(axis=(1, 3, 5))
    assert np.allclose(out, want)
    assert c.pack_moment(0, 4).shape == (4,) and np.allclose(c.pack_moment(0, 4), inner.reshape(-1, 4).mean(axis=0))
    assert len(c.pack_moment(0, 1)) == n ** 3 * 4


def test_pack_tracer_packs_negative_ids_only(oracle_port):
    """test_xtensor_packer3d.cpp:303-340"""
    from oracle import nixoracle as no
    c = no.Chunk(oracle_port, (4, 4, 4), 2, 2, ns=1, np_required=[16])
    xu = np.zeros((6, 7))
    xu[:, 0:3] = 1.5
    xu[:, 3] = np.arange(6)
    ids = np.array([5, -1, 7, -9, -3, 11], dtype=np.int64)
    xu[:, 6] = ids.view(np.float64)
    c.set_particles(0, xu)
    t = c.pack_tracer(0)
    assert t.shape == (3, 7)
    assert np.ascontiguousarray(t[:, 6]).view(np.int64).tolist() == [-1, -9, -3]
    assert t[:, 3].tolist() == [1.0, 3.0, 4.0]


# ---- XtensorParticle: the reference's own container tests (unittest/test_xtensor_particle.cpp) ----------------
@pytest.mark.parametrize("which", ["port", "ref"])
def test_create_particle_sizes(which):
    """test_xtensor_particle.cpp:210-232 "CreateParticle": Np = 1000 in an 8^3 chunk with one ghost layer ->
    Ng = 10^3, Np_total = ((Np + 128) / 128) * 128 slots (particle.hpp:146-153), pindex [Ng + 1], pcount
    [Ng + 1][8], everything zero."""
    if not no.available(which):
        pytest.skip(f"{which} library not built here")
    lib = no.load(which)
    c = no.Chunk(lib, (8, 8, 8), 1, 1, ns=1, np_required=[1000])
    assert c.ng(0) == 10 * 10 * 10
    assert c.np_total(0) == ((1000 + 128) // 128) * 128 == 1024
    assert c.xu(0).shape == (1024, 7) and c.xv(0).shape == (1024, 7) and c.gindex(0).shape == (1024,)
    assert c.pindex(0).shape == (1001,) and c.pcount(0).shape == (1001, 8)
    assert not c.xu(0).any() and not c.xv(0).any()


@pytest.mark.parametrize("npart", [100, 1000, 10000])
@pytest.mark.parametrize("dims", [(8, 8, 8), (16, 8, 16), (16, 16, 16)])
def test_sort_particle_3d(oracle_port, npart, dims):
    """test_xtensor_particle.cpp:339-367 "SortParticle3D" with what that test MEANT to check (it never sets Np, so
    the reference's own run is vacuous): positions uniform in the chunk plus one cell on every
    side, count(order 1) + sort, then every particle between pindex[ii] and pindex[ii + 1] lies in cell ii
    (check_sort3d, :190-207) and the out-of-bounds ones are gone."""
    rng = np.random.default_rng(npart + dims[0])
    nb = 1
    c = no.Chunk(oracle_port, dims, nb, 1, ns=1, np_required=[npart])
    x = np.zeros((npart, 7))
    for a in range(3):  # x, y, z <- dims[2], dims[1], dims[0]
        x[:, a] = rng.uniform(-1.0, dims[2 - a] + 1.0, npart)
    c.set_particles(0, x)
    c.count(0, 0, npart - 1, True, 1)
    c.sort(0)
    inside = np.ones(npart, dtype=bool)
    for a in range(3):
        inside &= (x[:, a] >= 0.0) & (x[:, a] < dims[2 - a])
    assert c.np(0) == int(inside.sum())
    xs, pin = c.particles(0), c.pindex(0)
    # bins as count() computes them for an odd order: half a cell below the chunk (xtensor_particle.hpp:328-348)
    ix = np.floor(xs[:, 0] + 0.5).astype(int)
    iy = np.floor(xs[:, 1] + 0.5).astype(int)
    iz = np.floor(xs[:, 2] + 0.5).astype(int)
    cell = (iz * (dims[1] + 1) + iy) * (dims[2] + 1) + ix
    assert (np.diff(cell) >= 0).all()
    i = np.arange(len(cell))
    assert (pin[cell] <= i).all() and (i < pin[cell + 1]).all()
    assert sorted(map(tuple, xs[:, :3])) == sorted(map(tuple, x[inside][:, :3]))


# ---- shift_weights: the reference's own vectors (test_interp.cpp:52-115, test_esirkepov.cpp:100-202) ----------
INTERP_WW = {1: [0.5, 0.5, 0.0], 2: [0.2, 0.6, 0.2, 0.0], 3: [0.1, 0.4, 0.4, 0.1, 0.0], 4: [0.1, 0.2, 0.4, 0.2, 0.1, 0.0]}
ESIRKEPOV_WW = {1: [0.0, 0.5, 0.5, 0.0], 2: [0.0, 0.2, 0.6, 0.2, 0.0], 3: [0.0, 0.1, 0.4, 0.4, 0.1, 0.0],
                4: [0.0, 0.1, 0.2, 0.4, 0.2, 0.1, 0.0]}


def _libs():
    return [no.load(w) for w in ("port", "ref") if no.available(w)]


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_interp_shift_weights(order):
    """test_interp.cpp:52-115: shift 0 leaves the (order+2)-wide window alone, shift 1 moves it up one slot and
    clears slot 0 (half-grid weights aligned to the common stencil, interp.hpp:149-160)."""
    for lib in _libs():
        for shift in (0, 1):
            ww = np.array(INTERP_WW[order])
            lib.nixo_interp_shift_weights(order, shift, ww.ctypes.data_as(PD))
            want = INTERP_WW[order] if shift == 0 else [0.0] + INTERP_WW[order][:-1]
            assert ww.tolist() == want


@pytest.mark.parametrize("order", [1, 2, 3, 4])
@pytest.mark.parametrize("shift", [(+1, -1, 0), (0, +1, +1), (+1, 0, -1), (-1, -1, -1), (0, 0, 0)])
def test_esirkepov_shift_weights(order, shift):
    """test_esirkepov.cpp:100-202: a particle that moved one cell down has its (order+3)-wide weights moved one slot
    to the left, one cell up one slot to the right (esirkepov.hpp:241-258); the padded window keeps its sum."""
    for lib in _libs():
        ss = np.array([ESIRKEPOV_WW[order]] * 3)
        sh = (C.c_int * 3)(*shift)
        lib.nixo_esirkepov_shift_weights(order, sh, ss.ctypes.data_as(PD))
        for d in range(3):
            w = ESIRKEPOV_WW[order]
            want = w if shift[d] == 0 else (w[1:] + [w[-1]] if shift[d] < 0 else [w[0]] + w[:-1])
            assert ss[d].tolist() == want, (order, shift, d)
            assert abs(ss[d].sum() - 1.0) < 1e-15


# ---- append_current3d / append_moment3d (test_primitives.cpp:1008-1100, 1236-1300, 1633-1657) -----------------
@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_append_current3d_adds_the_local_mesh(order):
    """test_primitives.cpp:1008-1100 via test_append_current3d_scalar (:1633-1657): five appends of a random
    (order+3)^3 x 4 mesh at (2, 2, 2) of a zeroed 16^3 array leave 5 x the mesh there -- and nothing anywhere else
    (the reference's check sums the error over the box only)."""
    n, N = order + 3, 16
    rng = np.random.default_rng(order)
    cur = rng.uniform(0, 1, (n, n, n, 4))
    out = []
    for lib in _libs():
        uj = np.zeros((N, N, N, 4))
        for _ in range(5):
            lib.nixo_append_current3d(order, uj.ctypes.data_as(PD), N, N, 2, 2, 2, cur.ctypes.data_as(PD))
        box = uj[2:2 + n, 2:2 + n, 2:2 + n]
        assert np.abs(box - 5 * cur).sum() <= 1e-14 * np.abs(box).sum()
        uj[2:2 + n, 2:2 + n, 2:2 + n] = 0
        assert not uj.any()
        out.append(box.copy())
    assert all(np.array_equal(out[0], o) for o in out)  # port == reference, bit for bit


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_append_moment3d_adds_the_local_mesh(order):
    """test_primitives.cpp:1236-1300: the same for the 14 moments of one species on the (order+1)^3 mesh."""
    n, N, ns = order + 1, 16, 2
    rng = np.random.default_rng(10 + order)
    mom = rng.uniform(0, 1, (n, n, n, 14))
    out = []
    for lib in _libs():
        um = np.zeros((N, N, N, ns, 14))
        for _ in range(5):
            lib.nixo_append_moment3d(order, um.ctypes.data_as(PD), N, N, ns, 3, 2, 4, 1, mom.ctypes.data_as(PD))
        box = um[3:3 + n, 2:2 + n, 4:4 + n, 1]
        assert np.abs(box - 5 * mom).sum() <= 1e-14 * np.abs(box).sum()
        assert not um[..., 0, :].any()
        um[3:3 + n, 2:2 + n, 4:4 + n, 1] = 0
        assert not um.any()
        out.append(box.copy())
    assert all(np.array_equal(out[0], o) for o in out)
