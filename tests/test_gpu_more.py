"""More GPU parity (through the C ABI): golden fixtures generated from the reference, the per-chunk
halo buffers in the reference's MpiBuffer byte layout, edge cases (empty / ragged chunks, every
particle out of bounds), error behaviour, and size-independent properties at sizes the oracle would
not finish in seconds (charge, continuity, conservation, sortedness)."""
import os

import numpy as np
import pytest

from nix_b200 import core
from nix_b200.synth import Problem
from oracle import nixoracle as no

from helpers import (assert_particles_equal, bits, gpu_domain, gpu_domain_from_golden, oracle_domain)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("order,tag", [(1, ""), (2, ""), (3, ""), (2, "_vay"), (2, "_hc")])
def test_golden_fixture_from_reference(gpu_lib, order, tag):
    """tests/golden/steps_order*.npz were written by the reference's own templates (oracle/_ref); the
    _vay / _hc fixtures use push_vay / push_higuera_cary instead of push_boris."""
    g = np.load(os.path.join(GOLD, f"steps_order{order}{tag}.npz"))
    gd = gpu_domain_from_golden(g, order, strict=True)
    ns = len(g["q"])
    for k in range(gd.nchunk):
        for s in range(ns):
            assert np.array_equal(bits(gd.get_particles(k, s)), bits(g[f"sorted_xu_{k}_{s}"]))
            assert np.array_equal(gd.get_pindex(k, s), g[f"sorted_pindex_{k}_{s}"])
    for _ in range(int(g["nstep"])):
        gd.step(float(g["delt"]))
    assert gd.check() == 0
    for k in range(gd.nchunk):
        assert np.array_equal(bits(gd.get_field(k)), bits(g[f"out_uf_{k}"]))
        ref = g[f"out_uj_{k}"]
        assert np.abs(gd.get_current(k) - ref).max() <= 1e-12 * np.abs(ref).max()
        for s in range(ns):
            assert np.array_equal(bits(gd.get_particles(k, s)), bits(g[f"out_xu_{k}_{s}"]))
            assert np.array_equal(gd.get_pindex(k, s), g[f"out_pindex_{k}_{s}"])
    gd.close()


@pytest.mark.parametrize("order,nb", [(1, 2), (2, 2), (3, 3), (2, 3)])
@pytest.mark.parametrize("mode", [no.MODE_FIELD, no.MODE_CURRENT])
def test_halo_buffers_are_the_reference_mpibuffer(oracle_port, gpu_lib, order, nb, mode):
    """nixb200_chunk_halo_pack produces the bytes Chunk::pack_bc_exchange leaves in the send buffer
    (layout chunk.cpp:257-286), for all 26 directions; unpack consumes the reference's buffer."""
    prob = Problem((2, 2, 2), (6, 8, 10), order, ppc=2, seed=5 + order, nb=nb)
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob)
    rng = np.random.default_rng(3)
    for k, c in enumerate(od.chunks):
        c.uj[...] = rng.normal(size=c.uj.shape)
        gd.set_current(k, c.uj)
    bs, ba = gd.halo_layout(mode)
    for k, c in enumerate(od.chunks):
        c.halo_pack(mode)
        assert np.array_equal(bs, c.bufsize(mode)) and np.array_equal(ba, c.bufaddr(mode))
        assert np.array_equal(gd.halo_pack(k, mode), c.sendbuf(mode)), f"send buffer of chunk {k}"
    # feed chunk 0 the buffers its neighbours produced, on both sides
    c0 = od.chunks[0]
    recv = np.zeros(int(ba[26] + bs[26]), dtype=np.uint8)
    for e in range(27):
        if e == 13:
            continue
        n = od.neighbor(0, e // 9, (e // 3) % 3, e % 3)
        sb = od.chunks[n].sendbuf(mode)
        recv[ba[e]:ba[e] + bs[e]] = sb[ba[26 - e]:ba[26 - e] + bs[26 - e]]
    c0.set_recv_sizes(mode, bs)
    c0.recvbuf(mode)[:] = recv
    c0.halo_unpack(mode)
    gd.halo_unpack(0, mode, recv)
    got = gd.get_field(0) if mode == no.MODE_FIELD else gd.get_current(0)
    ref = c0.uf if mode == no.MODE_FIELD else c0.uj
    assert np.array_equal(bits(got), bits(ref))
    gd.close()


def test_halo_unpack_skips_invalid_neighbours(oracle_port, gpu_lib):
    """MPI_PROC_NULL neighbours are skipped (xtensor_halo3d.hpp:57): only +x valid, as the
    reference's own halo tests do (test_xtensor_halo3d.cpp:39-49)."""
    prob = Problem((1, 1, 1), (4, 4, 4), 1, ppc=1, seed=2)
    gd = gpu_domain(prob)
    M = gd.M
    iz, iy, ix, c = np.meshgrid(*[np.arange(m) for m in M], np.arange(6), indexing="ij")
    pattern = (iz * 1e5 + iy * 1e3 + ix * 10 + c).astype(np.float64)  # test_xtensor_halo3d.cpp:71-78
    gd.set_field(0, pattern)
    buf = gd.halo_pack(0, no.MODE_FIELD)
    bs, ba = gd.halo_layout(no.MODE_FIELD)
    recv = np.zeros_like(buf)
    e = 9 * 1 + 3 * 1 + 2  # slot (1,1,2): from the +x neighbour, which sent in direction (1,1,0)
    d = 26 - e
    recv[ba[e]:ba[e] + bs[e]] = buf[ba[d]:ba[d] + bs[d]]
    valid = np.zeros(27, dtype=np.int32)
    valid[e] = 1
    gd.halo_unpack(0, no.MODE_FIELD, recv, valid)
    out = gd.get_field(0)
    nb, N = prob.nb, 4
    assert np.array_equal(out[nb:nb + N, nb:nb + N, nb + N:], pattern[nb:nb + N, nb:nb + N, nb:2 * nb])
    mask = np.ones(M + (6,), dtype=bool)
    mask[nb:nb + N, nb:nb + N, nb + N:] = False
    assert np.array_equal(out[mask], pattern[mask])  # nothing else touched
    gd.close()


def test_empty_and_ragged_chunks(oracle_port, gpu_lib):
    """Chunks with no particles, one particle, and very different counts; a species that is empty
    everywhere."""
    prob = Problem((2, 2, 2), (8, 8, 8), 2, ppc=6, seed=17, vth=(0.3, 0.05),
                   density=lambda c, cd: [0.0, 1.0 / 3072, 0.3, 2.5][(int(c[0]) * 2 + int(c[1])) % 4])
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, capacity_factor=3.0)
    assert 0 in [od.chunks[k].np(0) for k in range(8)] and 1 in [od.chunks[k].np(0) for k in range(8)]
    for step in range(3):
        od.step(0.5, 1.0)
        gd.step(0.5)
        assert gd.check() == 0
        assert_particles_equal(od, gd, f"ragged step {step}")
        for k, c in enumerate(od.chunks):
            s = np.abs(c.uj).max() or 1.0
            assert np.abs(gd.get_current(k) - c.uj).max() / s < 1e-12
    gd.close()
    # a completely empty domain steps without error
    empty = Problem((1, 1, 2), (8, 8, 8), 1, ppc=0, seed=1)
    ge = gpu_domain(empty)
    ge.step(0.5)
    assert ge.check() == 0 and ge.total_particles() == 0
    assert np.abs(ge.get_current(0)).max() == 0.0
    ge.close()


def test_all_particles_out_of_bounds_are_dropped(oracle_port, gpu_lib):
    """sort() keeps Np = pindex[Ng]: out-of-bounds particles disappear (xtensor_particle.hpp:320)."""
    prob = Problem((1, 1, 1), (8, 8, 8), 2, ppc=3, seed=4, oob_frac=1.0)
    od = oracle_domain(oracle_port, prob, fields=False)
    gd = gpu_domain(prob, fields=False)
    assert od.total_particles() == 0 == gd.total_particles()
    pc = gd.get_pcount(0, 0)
    assert pc[:-1].sum() == 0 and pc[-1].sum() == prob.ncell() * prob.ppc  # all in the row Ng
    gd.close()


def test_error_behaviour(gpu_lib):
    """int status + last_error text, no exception across the C boundary."""
    from nix_b200 import core
    prob = Problem((1, 1, 1), (8, 8, 8), 3, ppc=1)
    with pytest.raises(core.NixB200Error, match="boundary margin"):
        core.Domain(prob.cdims, prob.dims, 2, 3, prob.q, prob.m)  # order 3 needs nb >= 3
    with pytest.raises(core.NixB200Error, match="order"):
        core.Domain(prob.cdims, prob.dims, 3, 4, prob.q, prob.m)
    d = core.Domain(prob.cdims, prob.dims, 3, 3, prob.q, prob.m)
    with pytest.raises(core.NixB200Error, match="chunk index"):
        d.get_field(5)
    with pytest.raises(core.NixB200Error, match="species"):
        d.get_np(7)
    with pytest.raises(core.NixB200Error, match="set_ranks"):
        d._ck(d.lib.nixb200_domain_comm_init(d.h, b"\0" * 128))
    d.close()


def test_cfl_violation_is_flagged(gpu_lib):
    """A particle that moves more than one cell deposits nothing and raises NIXB200_ERR_CFL."""
    from nix_b200 import core
    prob = Problem((1, 1, 1), (8, 8, 8), 2, ppc=1, seed=9, vth=(0.01, 0.01))
    gd = gpu_domain(prob)
    gd.clear_current()
    gd.push_deposit(5.0)  # c*dt = 5 cells for relativistic particles is not reached at vth 0.01...
    assert gd.check() == 0
    fast = Problem((1, 1, 1), (8, 8, 8), 2, ppc=1, seed=9, vth=(30.0, 30.0))
    gf = gpu_domain(fast)
    gf.clear_current()
    gf.push_deposit(3.0)
    assert gf.check() & core.ERR_CFL
    gd.close()
    gf.close()


@pytest.mark.parametrize("order", [1, 2, 3])
def test_properties_at_scale(gpu_lib, order):
    """4x4x4 chunks of 16^3, 32 ppc x 2 species (8.4 M particles): properties that need no oracle.
       * particle number and total charge are conserved over steps with migration;
       * sum of rho over the interior after the J halo = total charge of the particles
         (test_esirkepov.cpp:1007-1010), at round-off;
       * discrete continuity: (rho_new - rho_old) + dt/dh * div J = 0 cell by cell
         (test_esirkepov.cpp:993-1028; rho = the rho slot of two successive deposits), at round-off;
       * sortedness: every chunk is cell-ordered, pindex brackets every cell, the sort is a permutation."""
    prob = Problem((4, 4, 4), (16, 16, 16), order, ppc=32, seed=100 + order, vth=(0.2, 0.04))
    gd = gpu_domain(prob, strict=False, capacity_factor=1.3)
    nb, N = prob.nb, 16
    ntot = gd.total_particles()
    assert ntot == prob.total_particles()
    qtot = float(sum(prob.q[s] * prob.ncell() * prob.ppc * prob.nchunk for s in range(prob.ns)))
    inner = (slice(nb, nb + N),) * 3

    def global_grid(which):
        out = np.zeros((64, 64, 64, 4))
        for k in range(gd.nchunk):
            cz, cy, cx = (int(v) * N for v in prob.coord[k])
            out[cz:cz + N, cy:cy + N, cx:cx + N] = which(k)[inner]
        return out

    dt = 0.5
    gd.step(dt)
    j0 = global_grid(gd.get_current)
    gd.step(dt)
    j1 = global_grid(gd.get_current)
    assert gd.check() == 0
    assert gd.total_particles() == ntot
    # rho slot: the deposit leaves rho of the NEW positions; sum = net charge (0 for e-/ion pairs),
    # so compare per species sign-free through |rho| scale
    scale = np.abs(j1[..., 0]).sum()
    assert abs(j1[..., 0].sum() - qtot) <= 1e-12 * scale
    # continuity between the two steps (periodic differences on the global grid)
    div = ((np.roll(j1[..., 1], -1, axis=2) - j1[..., 1]) + (np.roll(j1[..., 2], -1, axis=1) - j1[..., 2]) +
           (np.roll(j1[..., 3], -1, axis=0) - j1[..., 3]))
    res = (j1[..., 0] - j0[..., 0]) + dt / 1.0 * div
    assert np.abs(res).sum() <= 1e-12 * scale, f"continuity residual {np.abs(res).sum() / scale:.3e}"
    # the sort is a permutation (ids preserved) that leaves every chunk cell-ordered, with pindex
    # bracketing every cell (xtensor_particle.hpp:260-321); bins as count() computes them (:328-348).
    # (It is NOT idempotent in the reference either: the key cell*8 + ip%8 depends on the position.)
    is_odd = order % 2
    for k in (0, 21, 63):
        lo = np.array([int(v) * N for v in prob.coord[k]], dtype=np.float64)
        for s in range(prob.ns):
            before = gd.get_particles(k, s)
            pin = gd.get_pindex(k, s)
            iz = np.floor((before[:, 2] - (lo[0] - 0.5 * is_odd)) * 1.0).astype(np.int64)
            iy = np.floor((before[:, 1] - (lo[1] - 0.5 * is_odd)) * 1.0).astype(np.int64)
            ix = np.floor((before[:, 0] - (lo[2] - 0.5 * is_odd)) * 1.0).astype(np.int64)
            cell = (iz * (N + 1) + iy) * (N + 1) + ix
            assert np.all(np.diff(cell) >= 0), "particles are not cell-ordered"
            i = np.arange(len(cell))
            assert np.all(pin[cell] <= i) and np.all(i < pin[cell + 1])
            assert pin[-1] == len(cell)
    ids_before = [np.sort(bits(gd.get_particles(k, s)[:, 6])) for k in (0, 21, 63) for s in range(prob.ns)]
    gd.sort()
    ids_after = [np.sort(bits(gd.get_particles(k, s)[:, 6])) for k in (0, 21, 63) for s in range(prob.ns)]
    assert all(np.array_equal(a, b) for a, b in zip(ids_before, ids_after))
    assert gd.total_particles() == ntot
    gd.close()


class WeibelProblem(Problem):
    """BASELINE.json configs[3] in miniature: two counter-streaming electron species, u_z = +-0.5 c,
    cold (T = 0.001), 3rd-order shape."""

    def __init__(self, cdims, dims, ppc, seed=77):
        super().__init__(cdims, dims, 3, ppc=ppc, ns=2, seed=seed, vth=(0.03, 0.03), q=(-1.0, -1.0), m=(1.0, 1.0))

    def particles(self, k, s):
        xu = super().particles(k, s)
        xu[:, 5] += 0.5 if s == 0 else -0.5
        return xu


def test_weibel_counter_streaming_order3(oracle_port, gpu_lib):
    """Counter-streaming beams at order 3 against the oracle: every particle crosses a bin boundary every
    few steps along z (the mover path of k_deposit and the z-direction migration carry the load)."""
    prob = WeibelProblem((2, 2, 2), (8, 8, 8), ppc=8)
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True)
    for step in range(5):
        od.step(0.5, 1.0)
        gd.step(0.5)
        assert gd.check() == 0
        for k, c in enumerate(od.chunks):
            err = np.abs(gd.get_current(k) - c.uj).max() / np.abs(c.uj).max()
            assert err < 1e-12, f"step {step} J chunk {k}: {err:.2e}"
        assert_particles_equal(od, gd, f"weibel step {step}")
    gd.close()


def test_chunks_of_32_cubed_in_gilbert_order(oracle_port, gpu_lib):
    """BASELINE.json configs[2] in miniature: 32^3-cell chunks (several push tiles per axis, 33 bins per
    row) visited in space-filling-curve order, order 2, against the oracle over steps with migration."""
    prob = Problem((2, 2, 2), (32, 32, 32), 2, ppc=3, seed=321, vth=(0.3, 0.06))
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True)
    for step in range(2):
        od.step(0.5, 1.0)
        gd.step(0.5)
        assert gd.check() == 0
        for k, c in enumerate(od.chunks):
            err = np.abs(gd.get_current(k) - c.uj).max() / np.abs(c.uj).max()
            assert err < 1e-12, f"step {step} J chunk {k}: {err:.2e}"
            for s in range(prob.ns):
                assert np.array_equal(gd.get_pindex(k, s), c.pindex(s))
        assert_particles_equal(od, gd, f"32^3 step {step}")
    gd.close()


def test_overlapped_field_transfers(gpu_lib):
    """The pipelined host loop of bench.py's e2e leg (E/B up and J down on the domain's copy stream while
    the device migrates and sorts) gives the results of the plain step: particles bit for bit, J to
    round-off (the deposit adds with atomics), and the uploaded E/B arrives unchanged."""
    import torch
    from nix_b200 import core
    prob = Problem((2, 2, 2), (8, 8, 8), 2, ppc=16, seed=9, vth=(0.3, 0.05))
    ga = gpu_domain(prob, strict=True)
    gb = gpu_domain(prob, strict=True)
    cells = int(np.prod(prob.M))
    uf_host = torch.empty((gb.nchunk, cells, 6), dtype=torch.float64, pin_memory=True)
    uj_host = torch.zeros((gb.nchunk, cells, 4), dtype=torch.float64, pin_memory=True)
    for k in range(gb.nchunk):
        uf_host.numpy()[k] = ga.get_field(k).reshape(cells, 6)
    gb.field_upload_overlapped(core.FIELD_UF, uf_host.data_ptr())
    for step in range(3):
        ga.step(0.5)
        gb.clear_current()
        gb.push_deposit(0.5)
        gb.exchange_current()
        gb.field_download_overlapped(core.FIELD_UJ, uj_host.data_ptr())
        gb.exchange_field()
        gb.migrate_sort()
        gb.copy_synchronize()
        gb.field_upload_overlapped(core.FIELD_UF, uf_host.data_ptr())
        assert ga.check() == 0 and gb.check() == 0
        for k in range(ga.nchunk):
            ja = ga.get_current(k)
            jb = uj_host.numpy()[k].reshape(ja.shape)
            assert np.abs(ja - jb).max() <= 1e-12 * np.abs(ja).max(), f"step {step} chunk {k}"
            assert np.array_equal(bits(ga.get_field(k)), bits(gb.get_field(k)))
            for s in range(prob.ns):
                assert np.array_equal(bits(ga.get_particles(k, s)), bits(gb.get_particles(k, s)))
    ga.close()
    gb.close()


def test_interior_transfers(gpu_lib):
    """Interior-only overlapped transfers (the host-side field solver's view): the J that comes down is the
    interior of get_current, and interior E/B uploaded + E/B halo reproduces the full array."""
    import torch
    from nix_b200 import core
    prob = Problem((2, 2, 2), (8, 6, 10), 2, ppc=8, seed=19, vth=(0.3, 0.05))
    gd = gpu_domain(prob, strict=True)
    nb, (nz, ny, nx) = prob.nb, prob.dims
    inner = (slice(nb, nb + nz), slice(nb, nb + ny), slice(nb, nb + nx))
    full0 = [gd.get_field(k) for k in range(gd.nchunk)]
    ufi = torch.empty((gd.nchunk, nz, ny, nx, 6), dtype=torch.float64, pin_memory=True)
    uji = torch.zeros((gd.nchunk, nz, ny, nx, 4), dtype=torch.float64, pin_memory=True)
    for k in range(gd.nchunk):
        ufi.numpy()[k] = full0[k][inner]
        gd.set_field(k, np.full_like(full0[k], 7.0))  # wipe: interior and ghosts
    gd.interior_upload_overlapped(core.FIELD_UF, ufi.data_ptr())
    gd.exchange_field()
    for k in range(gd.nchunk):
        assert np.array_equal(bits(gd.get_field(k)), bits(full0[k])), f"E/B chunk {k}"
    gd.clear_current()
    gd.push_deposit(0.5)
    gd.exchange_current()
    gd.interior_download_overlapped(core.FIELD_UJ, uji.data_ptr())
    gd.migrate_sort()
    gd.copy_synchronize()
    assert gd.check() == 0
    for k in range(gd.nchunk):
        assert np.array_equal(bits(uji.numpy()[k]), bits(gd.get_current(k)[inner])), f"J chunk {k}"
    gd.close()


# ---- particle storage: growth and overflow (sort.cu:405, push_deposit.cu:814) ----------
def _stream_problem(n=40000):
    """Two chunks of (8, 8, 2) cells side by side in x; every particle flies in +x at ~c, so c*dt/dx of a
    chunk's particles cross into the neighbour every step."""
    prob = Problem((1, 1, 2), (8, 8, 2), 2, ppc=1, ns=1, seed=5, vth=(0.0,), amp=0.0, b0=0.0)

    def particles(k, s, prob=prob):
        rng = np.random.default_rng([99, k])
        xu = np.zeros((n, 7))
        xu[:, 0] = 2.0 * k + rng.uniform(0.0, 2.0, n) * (1 - 1e-12)
        xu[:, 1] = rng.uniform(0.0, 8.0, n) * (1 - 1e-12)
        xu[:, 2] = rng.uniform(0.0, 8.0, n) * (1 - 1e-12)
        xu[:, 3] = 50.0  # u_x: v = 0.9998 c
        xu[:, 6] = (np.arange(n, dtype=np.int64) + (np.int64(k) << 32)).view(np.float64)
        return xu
    prob.particles = particles
    return prob


def test_migration_buffers_regrow_before_they_fill(oracle_port, gpu_lib):
    """5 % of the particles leave in step 1 (the buffers hold 6 %): the next step regrows the buffers, so the
    15 % and then 45 % that leave in the following steps fit.  Particles stay bit-exact with the oracle."""
    prob = _stream_problem()
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True, capacity_factor=1.0)
    cap0, lcap0 = gd.capacity(0)
    assert lcap0 < 0.1 * 2 * 40000
    for dt in (0.1, 0.3, 0.9, 0.9):
        od.step(dt, 1.0)
        gd.step(dt)
        assert gd.check() == 0
        assert_particles_equal(od, gd, f"dt {dt}")
    cap1, lcap1 = gd.capacity(0)
    assert lcap1 > lcap0, "the migration buffers did not grow"
    gd.close()


def test_migration_overflow_is_an_error_not_a_memory_fault(oracle_port, gpu_lib):
    """45 % leave in the very first step, more than the buffers hold and with nothing to warn the host:
    ERR_CAPACITY, no out-of-bounds access, and the next step refuses to run; with room reserved up front
    the same run is bit-exact."""
    prob = _stream_problem()
    gd = gpu_domain(prob, strict=True, capacity_factor=1.0)
    gd.step(0.9)
    assert gd.check() & core.ERR_CAPACITY
    with pytest.raises(core.NixB200Error, match="overflowed"):
        gd.step(0.9)
    gd.close()
    od = oracle_domain(oracle_port, prob)
    gd = gpu_domain(prob, strict=True, capacity_factor=1.0)
    gd.reserve(0, nmove=60000)
    for _ in range(2):
        od.step(0.9, 1.0)
        gd.step(0.9)
        assert gd.check() == 0
        assert_particles_equal(od, gd, "reserved")
    gd.close()
