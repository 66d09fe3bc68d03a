"""Known answers for the oracle's Yee field update (oracle/field_solver.c).

The reference ships no field solver, so no fixture of its own can pin this one ("parity unpinned by the
reference"); the oracle is pinned by the properties the scheme must have on the lattice the reference
does fix (xtensor_packer3d.hpp:279-302, esirkepov.hpp:177-237): exact discrete plane-wave eigenmodes with
the Yee dispersion relation, div B = 0, and Gauss's law preserved to round-off together with the
reference's Esirkepov deposit."""
import numpy as np
import pytest

from nix_b200.synth import Problem
from oracle import nixoracle as no

from helpers import oracle_domain


def _vacuum_domain(lib, cdims, dims, delh, nb=2):
    prob = Problem(cdims, dims, 2, ppc=0, ns=1, nb=nb, delh=delh)
    d = no.Domain(lib, prob.cdims, prob.dims, prob.nb, 2, 1, prob.q, prob.m, prob.coord, 1, delh=delh)
    return prob, d


def _em_vacuum_step(d, dt, cc):
    d.push_bfd(0.5 * dt, cc, ext=1)
    d.push_efd(dt, cc, 0.0)
    d.exchange(no.MODE_FIELD)
    d.push_bfd(0.5 * dt, cc, ext=0)
    d.exchange(no.MODE_FIELD)


@pytest.mark.parametrize("cdims", [(1, 1, 1), (1, 2, 2)])
def test_plane_wave_eigenmode(oracle_port, cdims):
    """Ez = A sin(kx x + ky y) on its own lattice sites with B = 0 is a standing eigenmode:
    E^n = cos(n w dt) E^0 with sin^2(w dt / 2) = (c dt)^2 sum_a (sin(k_a d_a / 2) / d_a)^2."""
    dims, delh, cc, dt = (4, 12, 16), (1.0, 0.8, 1.25), 1.0, 0.4
    prob, d = _vacuum_domain(oracle_port, cdims, dims, delh)
    Ly, Lx = cdims[1] * dims[1] * delh[1], cdims[2] * dims[2] * delh[2]
    ky, kx = 2 * np.pi * 2 / Ly, 2 * np.pi * 3 / Lx
    nb = prob.nb

    def mode(k):
        c = prob.coord[k]
        iy = np.arange(-nb, dims[1] + nb) + c[1] * dims[1]
        ix = np.arange(-nb, dims[2] + nb) + c[2] * dims[2]
        yc, xc = (iy + 0.5) * delh[1], (ix + 0.5) * delh[2]  # Ez sits at (edge, centre, centre)
        return np.broadcast_to(np.sin(ky * yc[:, None] + kx * xc[None, :]), (dims[0] + 2 * nb,) + (len(iy), len(ix)))
    for k, c in enumerate(d.chunks):
        c.uf[...] = 0.0
        c.uf[..., 2] = 0.7 * mode(k)
    s2 = (cc * dt) ** 2 * ((np.sin(ky * delh[1] / 2) / delh[1]) ** 2 + (np.sin(kx * delh[2] / 2) / delh[2]) ** 2)
    w = 2 * np.arcsin(np.sqrt(s2)) / dt
    I = (slice(nb, nb + dims[0]), slice(nb, nb + dims[1]), slice(nb, nb + dims[2]))
    for n in range(1, 25):
        _em_vacuum_step(d, dt, cc)
        for k, c in enumerate(d.chunks):
            want = 0.7 * np.cos(n * w * dt) * mode(k)
            assert np.abs(c.uf[..., 2][I] - want[I]).max() < 2e-13, f"step {n}"
            assert np.abs(c.uf[..., 0:2][I]).max() < 1e-14


def test_div_b_stays_zero(oracle_port):
    cdims, dims, delh = (2, 1, 2), (6, 8, 10), (0.9, 1.0, 1.1)
    prob, d = _vacuum_domain(oracle_port, cdims, dims, delh)
    rng = np.random.default_rng(3)
    for c in d.chunks:
        c.uf[...] = 0.0
        c.uf[..., 0:3] = rng.uniform(-1, 1, c.uf.shape[:3] + (3,))
    d.exchange(no.MODE_FIELD)
    for _ in range(5):
        _em_vacuum_step(d, 0.3, 1.0)
    nb = prob.nb
    for c in d.chunks:
        u = c.uf
        s = slice(nb, -nb)
        m = slice(nb - 1, -nb - 1)
        div = ((u[s, s, s, 3] - u[s, s, m, 3]) / delh[2] + (u[s, s, s, 4] - u[s, m, s, 4]) / delh[1]
               + (u[s, s, s, 5] - u[m, s, s, 5]) / delh[0])
        assert np.abs(u[..., 3:6]).max() > 0.1
        assert np.abs(div).max() < 1e-14


def gauss_residual(uf, uj, nb, delh, cfj):
    s = slice(nb, -nb)
    p = slice(nb + 1, -nb + 1 if nb > 1 else None)
    div = ((uf[s, s, p, 0] - uf[s, s, s, 0]) / delh[2] + (uf[s, p, s, 1] - uf[s, s, s, 1]) / delh[1]
           + (uf[p, s, s, 2] - uf[s, s, s, 2]) / delh[0])
    return div - cfj * uj[s, s, s, 0]


@pytest.mark.parametrize("order", [1, 2, 3])
def test_gauss_law_is_preserved_with_the_esirkepov_deposit(oracle_port, order):
    """div E - cfj rho does not change from step to step (to round-off): the field update inherits the
    charge conservation of esirkepov::deposit3d (test_esirkepov.cpp:993-1106) on the staggered lattice."""
    prob = Problem((2, 2, 1), (8, 8, 8), order, ppc=6, seed=41 + order, vth=(0.3, 0.05))
    d = oracle_domain(oracle_port, prob)
    cfj = 0.37
    res = []
    for step in range(4):
        d.step_em(0.5, 1.0, cfj)
        res.append([gauss_residual(c.uf, c.uj, prob.nb, prob.delh, cfj).copy() for c in d.chunks])
    scale = max(np.abs(c.uj[..., 0]).max() for c in d.chunks) * cfj
    for step in range(1, 4):
        for k in range(prob.nchunk):
            drift = np.abs(res[step][k] - res[step - 1][k]).max() / scale
            assert drift < 5e-14, f"order {order} step {step} chunk {k}: {drift:.2e}"
    assert np.abs(res[0][0]).max() / scale > 1e-3  # (the residual itself is not small: the fields are random)
