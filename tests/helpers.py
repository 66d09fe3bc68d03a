"""Shared helpers for the parity tests: load one synthetic Problem into the oracle and the GPU."""
import numpy as np

from oracle import nixoracle as no


def oracle_domain(lib, prob, sort=True, fields=True):
    d = no.Domain(lib, prob.cdims, prob.dims, prob.nb, prob.order, prob.ns, prob.q, prob.m, prob.coord,
                  prob.ncell() * prob.ppc, delh=prob.delh)
    for k, c in enumerate(d.chunks):
        if fields:
            c.uf[...] = prob.field(k)
        for s in range(prob.ns):
            c.set_particles(s, prob.particles(k, s))
    if fields:
        d.exchange(no.MODE_FIELD)
    if sort:
        d.sort_only()
    return d


def gpu_domain(prob, strict=True, sort=True, fields=True, **kw):
    from nix_b200 import core
    d = core.Domain(prob.cdims, prob.dims, prob.nb, prob.order, prob.q, prob.m, delh=prob.delh, coord=prob.coord,
                    strict_fp=strict, **kw)
    if fields:
        for k in range(d.nchunk):
            d.set_field(k, prob.field(k))
        d.exchange_field()
    for s in range(prob.ns):
        d.set_particles(s, [prob.particles(k, s) for k in range(d.nchunk)])
    if sort:
        d.sort()
    return d


def bits(a):
    return np.ascontiguousarray(a).view(np.int64)


def assert_particles_equal(od, gd, what=""):
    for k, c in enumerate(od.chunks):
        for s in range(od.ns):
            ref = c.particles(s)
            got = gd.get_particles(k, s)
            assert got.shape == ref.shape, f"{what} chunk {k} species {s}: Np {got.shape[0]} != {ref.shape[0]}"
            assert np.array_equal(bits(got), bits(ref)), f"{what} chunk {k} species {s}: particles differ"


def ref_pcount_before_sort(c, s):
    """The reference turns pcount into end cursors inside sort(); undo that (xtensor_particle.hpp:265-313):
    after sort pcount.flat[j] = end address of bin j = start of bin j+1."""
    pc = c.pcount(s).reshape(-1).astype(np.int64)
    cnt = np.diff(np.concatenate([[0], pc]))
    return cnt.reshape(-1, 8).astype(np.int32)


def load_golden_domain(lib, g, order):
    """Build an oracle Domain from a tests/golden/steps_order*.npz fixture (inputs are stored in it)."""
    cdims, dims, nb = tuple(g["cdims"]), tuple(g["dims"]), int(g["nb"])
    q, m = g["q"], g["m"]
    ns = len(q)
    nchunk = int(np.prod(cdims))
    npr = max(len(g[f"in_xu_{k}_{s}"]) for k in range(nchunk) for s in range(ns))
    dom = no.Domain(lib, cdims, dims, nb, order, ns, q, m, g["coord"], npr)
    for k, c in enumerate(dom.chunks):
        c.uf[...] = g[f"in_uf_{k}"]
        for s in range(ns):
            c.set_particles(s, g[f"in_xu_{k}_{s}"])
    dom.exchange(no.MODE_FIELD)
    dom.sort_only()
    return dom, ns


def gpu_domain_from_golden(g, order, strict=True):
    from nix_b200 import core
    cdims, dims, nb = tuple(int(v) for v in g["cdims"]), tuple(int(v) for v in g["dims"]), int(g["nb"])
    q, m = g["q"], g["m"]
    pusher = int(g["pusher"]) if "pusher" in g.files else 0
    d = core.Domain(cdims, dims, nb, order, q, m, coord=g["coord"], strict_fp=strict, pusher=pusher)
    for k in range(d.nchunk):
        d.set_field(k, g[f"in_uf_{k}"])
    d.exchange_field()
    for s in range(len(q)):
        d.set_particles(s, [g[f"in_xu_{k}_{s}"] for k in range(d.nchunk)])
    d.sort()
    return d
