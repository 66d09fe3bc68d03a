"""GPU parity of the Yee field update (csrc/fdtd.cu) against oracle/field_solver.c:
the two kernels bit for bit, the composed electromagnetic step within the deposit's 1e-12, Gauss's law and
div B at round-off at a size the oracle is not run at, and the energy diagnostic."""
import numpy as np
import pytest

from nix_b200.synth import Problem
from oracle import nixoracle as no

from helpers import gpu_domain, oracle_domain
from test_field_solver import gauss_residual

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order,cdims,dims,delh", [(2, (2, 2, 2), (8, 8, 8), (1.0, 1.0, 1.0)),
                                                   (3, (1, 2, 3), (6, 8, 10), (0.9, 1.0, 1.1))])
def test_fdtd_kernels_bit_exact(oracle_port, gpu_lib, order, cdims, dims, delh):
    prob = Problem(cdims, dims, order, ppc=1, seed=7, delh=delh)
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True, cc=0.8)
    rng = np.random.default_rng(11)
    for k, c in enumerate(od.chunks):
        c.uj[...] = rng.uniform(-1, 1, c.uj.shape)
        gd.set_current(k, c.uj)
    for ext, dt in ((1, 0.2), (0, 0.35)):
        od.push_bfd(dt, 0.8, ext)
        gd.push_bfd(dt, ext)
        for k, c in enumerate(od.chunks):
            assert np.array_equal(gd.get_field(k), c.uf), f"push_bfd ext {ext} chunk {k}"
    od.push_efd(0.4, 0.8, 1.7)
    gd.push_efd(0.4, 1.7)
    for k, c in enumerate(od.chunks):
        assert np.array_equal(gd.get_field(k), c.uf), f"push_efd chunk {k}"
    e = gd.field_energy()
    assert np.allclose(e, od.field_energy(), rtol=1e-13, atol=0)
    gd.close()


def _by_id(xu):
    ids = np.ascontiguousarray(xu[:, 6]).view(np.int64)
    o = np.argsort(ids, kind="stable")
    return ids[o], xu[o]


@pytest.mark.parametrize("order", [1, 2, 3])
def test_electromagnetic_step_against_the_oracle(oracle_port, gpu_lib, order):
    """step_em over 4 steps with migration.  J agrees to 1e-12 of its maximum (summation order), hence E/B
    and from the second step on the particles agree to that level instead of bit for bit."""
    prob = Problem((2, 2, 2), (8, 8, 8), order, ppc=8, seed=51 + order, vth=(0.3, 0.06))
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True)
    for step in range(4):
        od.step_em(0.5, 1.0, 0.5)
        gd.step_em(0.5, 0.5)
        assert gd.check() == 0
        fs = max(np.abs(c.uf).max() for c in od.chunks)
        for k, c in enumerate(od.chunks):
            assert np.abs(gd.get_field(k) - c.uf).max() / fs < 1e-12, f"step {step} E/B chunk {k}"
            assert np.abs(gd.get_current(k) - c.uj).max() / np.abs(c.uj).max() < 1e-12, f"step {step} J chunk {k}"
            for s in range(prob.ns):
                rid, rx = _by_id(c.particles(s))
                gid, gx = _by_id(gd.get_particles(k, s))
                assert np.array_equal(rid, gid), f"step {step} chunk {k} species {s}: different particles"
                assert np.abs(gx[:, :6] - rx[:, :6]).max() < 1e-11
    gd.close()


def test_gauss_law_and_div_b_at_scale(gpu_lib):
    """4x4x4 chunks of 16^3, 16 ppc x 2 species, order 2, contracted arithmetic: over 5 electromagnetic steps
    div E - cfj rho does not drift and div B stays zero (both to round-off), particles are conserved."""
    prob = Problem((4, 4, 4), (16, 16, 16), 2, ppc=16, seed=99, vth=(0.1, 0.02))
    gd = gpu_domain(prob, strict=False)
    cfj, nb = 0.01, prob.nb  # electron plasma frequency sqrt(cfj ppc) = 0.4
    # start from fields with div B = 0: keep the random E, zero B
    for k in range(gd.nchunk):
        uf = gd.get_field(k)
        uf[..., 3:6] = 0.0
        gd.set_field(k, uf)
    ntot = gd.total_particles()
    prev = None
    for step in range(5):
        gd.step_em(0.5, cfj)
        assert gd.check() == 0
        res, scale, divb = [], 0.0, 0.0
        for k in range(gd.nchunk):
            uf, uj = gd.get_field(k), gd.get_current(k)
            res.append(gauss_residual(uf, uj, nb, prob.delh, cfj))
            scale = max(scale, np.abs(uj[..., 0]).max())
            s, m = slice(nb, -nb), slice(nb - 1, -nb - 1)
            divb = max(divb, np.abs((uf[s, s, s, 3] - uf[s, s, m, 3]) + (uf[s, s, s, 4] - uf[s, m, s, 4])
                                    + (uf[s, s, s, 5] - uf[m, s, s, 5])).max())
        assert divb < 1e-13
        if prev is not None:
            drift = max(np.abs(a - b).max() for a, b in zip(res, prev)) / scale
            assert drift < 1e-12, f"step {step}: Gauss residual drifted by {drift:.2e}"
        prev = res
    assert gd.total_particles() == ntot
    gd.close()


def test_history_read_back_without_synchronising(gpu_lib):
    """nixb200_domain_history_async: field energies and per-chunk particle counts of every step arrive in pinned
    host memory in stream order; after one synchronise they equal what the blocking calls return."""
    import torch
    prob = Problem((2, 2, 2), (8, 8, 8), 2, ppc=6, seed=3, vth=(0.3, 0.05))
    gd = gpu_domain(prob, strict=True)
    steps = 3
    e = torch.zeros((steps, gd.nchunk, 2), dtype=torch.float64, pin_memory=True)
    n = torch.zeros((steps, prob.ns, gd.nchunk), dtype=torch.int64, pin_memory=True)
    for k in range(steps):
        gd.step_em(0.5, 0.05)
        gd.history_async(e[k].data_ptr(), n[k].data_ptr())
    gd.synchronize()
    assert np.array_equal(e[-1].numpy(), gd.field_energy())
    for s in range(prob.ns):
        assert np.array_equal(n[-1, s].numpy(), gd.get_np(s))
    assert (n.sum(dim=(1, 2)) == prob.total_particles()).all()
    assert (e[0] != e[1]).any()
    gd.close()
