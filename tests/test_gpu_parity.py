"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical
seeded inputs.  Bit-exact for counts / permutations / positions / momenta (strict_fp), <= 1e-12 of
the field maximum for J (summation order differs), continuity at round-off."""
import numpy as np
import pytest

from nix_b200.synth import Problem
from oracle import nixoracle as no

from helpers import (assert_particles_equal, bits, gpu_domain, oracle_domain, ref_pcount_before_sort)

pytestmark = pytest.mark.gpu

JTOL = 1e-12


def rel_err(a, b):
    s = np.abs(b).max()
    return np.abs(a - b).max() / (s if s > 0 else 1.0)


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("cdims", [(1, 1, 1), (2, 2, 2)])
def test_sort_bit_exact(oracle_port, gpu_lib, order, cdims):
    prob = Problem(cdims, (8, 8, 8), order, ppc=10, oob_frac=0.05, seed=11 + order)
    od = oracle_domain(oracle_port, prob, sort=False, fields=False)
    gd = gpu_domain(prob, sort=False, fields=False)
    # counts as count() leaves them
    for k, c in enumerate(od.chunks):
        for s in range(prob.ns):
            c.count(s, 0, c.np(s) - 1, True)
    gd.sort()
    for k, c in enumerate(od.chunks):
        for s in range(prob.ns):
            assert np.array_equal(gd.get_pcount(k, s), c.pcount(s)), f"pcount chunk {k} species {s}"
            c.sort(s)
            assert np.array_equal(gd.get_pindex(k, s), c.pindex(s)), f"pindex chunk {k} species {s}"
    assert_particles_equal(od, gd, "sort")
    assert gd.check() == 0


@pytest.mark.parametrize("order", [1, 2, 3])
def test_push_deposit_strict(oracle_port, gpu_lib, order):
    prob = Problem((2, 2, 2), (8, 8, 8), order, ppc=12, seed=21 + order, vth=(0.3, 0.05))
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True)
    assert_particles_equal(od, gd, "initial sort")
    od.clear_current()
    od.push_deposit(0.5, 1.0)
    gd.clear_current()
    gd.push_deposit(0.5)
    assert gd.check() == 0
    assert_particles_equal(od, gd, "push (strict)")
    for k, c in enumerate(od.chunks):
        assert rel_err(gd.get_current(k), c.uj) < JTOL, f"J chunk {k}"


@pytest.mark.parametrize("order", [1, 2, 3])
def test_push_deposit_fast(oracle_port, gpu_lib, order):
    prob = Problem((2, 2, 2), (8, 8, 8), order, ppc=12, seed=31 + order, vth=(0.3, 0.05))
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=False)
    od.clear_current()
    od.push_deposit(0.5, 1.0)
    gd.clear_current()
    gd.push_deposit(0.5)
    assert gd.check() == 0
    for k, c in enumerate(od.chunks):
        for s in range(prob.ns):
            ref, got = c.particles(s), gd.get_particles(k, s)
            assert got.shape == ref.shape
            assert np.array_equal(bits(got[:, 6]), bits(ref[:, 6]))
            scale = np.maximum(np.abs(ref[:, :6]), 1.0)
            assert (np.abs(got[:, :6] - ref[:, :6]) / scale).max() < 1e-12
        assert rel_err(gd.get_current(k), c.uj) < JTOL


@pytest.mark.parametrize("order", [1, 2, 3])
def test_halo_exchange(oracle_port, gpu_lib, order):
    prob = Problem((2, 3, 2), (8, 6, 10), order, ppc=4, seed=41 + order, vth=(0.3, 0.05))
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob)
    # field ghosts after the exchange done by the loaders
    for k, c in enumerate(od.chunks):
        assert np.array_equal(gd.get_field(k), c.uf), f"field chunk {k}"
    od.clear_current()
    od.push_deposit(0.5, 1.0)
    # feed the oracle's J (before exchange) to the GPU so the exchange itself is compared bit for bit
    for k, c in enumerate(od.chunks):
        gd.set_current(k, c.uj)
    od.exchange(no.MODE_CURRENT)
    gd.exchange_current()
    for k, c in enumerate(od.chunks):
        assert np.array_equal(gd.get_current(k), c.uj), f"current chunk {k}"


@pytest.mark.parametrize("order,cdims,dims", [(1, (2, 2, 2), (8, 8, 8)), (2, (2, 2, 2), (8, 8, 8)),
                                              (3, (2, 2, 2), (8, 8, 8)), (2, (1, 1, 1), (12, 12, 12)),
                                              (2, (3, 2, 4), (6, 8, 5))])
def test_full_steps_strict(oracle_port, gpu_lib, order, cdims, dims):
    prob = Problem(cdims, dims, order, ppc=8, seed=51 + order, vth=(0.35, 0.08))
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True)
    ntot = od.total_particles()
    for step in range(4):
        od.step(0.5, 1.0)
        gd.step(0.5)
        assert gd.check() == 0
        for k, c in enumerate(od.chunks):
            assert rel_err(gd.get_current(k), c.uj) < JTOL, f"step {step} J chunk {k}"
            assert np.array_equal(gd.get_field(k), c.uf)
            for s in range(prob.ns):
                assert np.array_equal(gd.get_pindex(k, s), c.pindex(s)), f"step {step} pindex {k} {s}"
                assert np.array_equal(gd.get_pcount(k, s), ref_pcount_before_sort(c, s)), f"step {step} pcount {k} {s}"
        assert_particles_equal(od, gd, f"step {step}")
        assert gd.total_particles() == ntot == od.total_particles()


@pytest.mark.parametrize("pusher", [1, 2])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_other_pushers_strict(oracle_port, gpu_lib, order, pusher):
    """push_vay / push_higuera_cary (primitives.hpp:193-253) instead of push_boris: full steps with
    migration, particles bit for bit, J to round-off."""
    prob = Problem((2, 2, 2), (8, 8, 8), order, ppc=8, seed=61 + order, vth=(0.35, 0.08))
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True, pusher=pusher)
    oracle_port.nixo_set_pusher(pusher)
    try:
        for step in range(3):
            od.step(0.5, 1.0)
            gd.step(0.5)
            assert gd.check() == 0
            for k, c in enumerate(od.chunks):
                assert rel_err(gd.get_current(k), c.uj) < JTOL, f"step {step} J chunk {k}"
            assert_particles_equal(od, gd, f"pusher {pusher} step {step}")
    finally:
        oracle_port.nixo_set_pusher(0)
        gd.close()


@pytest.mark.parametrize("pusher", [1, 2])
def test_other_pushers_fast(oracle_port, gpu_lib, pusher):
    """Contracted (FMA) arithmetic: <= 1e-12 relative on positions and momenta after one push."""
    prob = Problem((2, 2, 2), (8, 8, 8), 2, ppc=12, seed=33, vth=(0.3, 0.05))
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=False, pusher=pusher)
    oracle_port.nixo_set_pusher(pusher)
    try:
        od.clear_current()
        od.push_deposit(0.5, 1.0)
    finally:
        oracle_port.nixo_set_pusher(0)
    gd.clear_current()
    gd.push_deposit(0.5)
    assert gd.check() == 0
    for k, c in enumerate(od.chunks):
        for s in range(prob.ns):
            ref, got = c.particles(s), gd.get_particles(k, s)
            assert got.shape == ref.shape
            scale = np.maximum(np.abs(ref[:, :6]), 1.0)
            assert (np.abs(got[:, :6] - ref[:, :6]) / scale).max() < 1e-12
        assert rel_err(gd.get_current(k), c.uj) < JTOL
    gd.close()


@pytest.mark.parametrize("order", [1, 2, 3])
def test_anisotropic_cells_and_c(oracle_port, gpu_lib, order):
    """delz != dely != delx and c != 1 (chunk.cpp:210-237): full steps with migration, strict."""
    prob = Problem((2, 2, 1), (8, 6, 10), order, ppc=6, seed=95 + order, vth=(0.5, 0.1), delh=(0.5, 1.25, 2.0))
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True, cc=2.0)
    for step in range(3):
        od.step(0.2, 2.0)
        gd.step(0.2)
        assert gd.check() == 0
        for k, c in enumerate(od.chunks):
            assert rel_err(gd.get_current(k), c.uj) < JTOL, f"step {step} J chunk {k}"
            for s in range(prob.ns):
                assert np.array_equal(gd.get_pindex(k, s), c.pindex(s))
        assert_particles_equal(od, gd, f"anisotropic step {step}")
    gd.close()


@pytest.mark.parametrize("ns", [1, 3])
def test_one_and_three_species(oracle_port, gpu_lib, ns):
    """Species count other than two (the J of all species accumulates into one array), nb larger than the
    stencil needs."""
    prob = Problem((2, 1, 2), (8, 8, 8), 2, ppc=6, ns=ns, seed=77 + ns, vth=(0.3, 0.05, 0.02)[:ns],
                   q=(-1.0, 1.0, 2.0)[:ns], m=(1.0, 25.0, 100.0)[:ns], nb=3)
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True)
    for step in range(3):
        od.step(0.5, 1.0)
        gd.step(0.5)
        assert gd.check() == 0
        for k, c in enumerate(od.chunks):
            assert rel_err(gd.get_current(k), c.uj) < JTOL, f"step {step} J chunk {k}"
        assert_particles_equal(od, gd, f"ns={ns} step {step}")
    gd.close()
