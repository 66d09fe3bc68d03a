"""Seeded synthetic inputs for the per-chunk PIC step.

A periodic box of Cz*Cy*Cx chunks, thermal electron-ion plasma: positions uniform in each chunk,
momenta Gaussian with a per-species thermal spread, E/B uniform in [-amp, amp] plus a guide field
B0 along z.  Everything is drawn from numpy's PCG64 with explicit seeds so that the CPU oracle and
the GPU run the same arrays.
"""
import numpy as np

from .sfc import chunk_coords


def snake_order(cdims):
    """Boustrophedon chunk order (consecutive ids are face neighbours).  NOT the reference's curve: kept
    only as a second, different id -> coordinate table for tests (the hot path takes any table)."""
    cz, cy, cx = cdims
    out = []
    for z in range(cz):
        ys = range(cy) if z % 2 == 0 else range(cy - 1, -1, -1)
        for iy, y in enumerate(ys):
            flip = (iy + z * cy) % 2 == 1
            xs = range(cx - 1, -1, -1) if flip else range(cx)
            for x in xs:
                out.append((z, y, x))
    return np.array(out, dtype=np.int32)


class Problem:
    def __init__(self, cdims, dims, order, ppc, ns=2, seed=1234, vth=(0.1, 0.02), amp=0.01, b0=0.1,
                 q=(-1.0, 1.0), m=(1.0, 25.0), coord=None, nb=None, delh=(1.0, 1.0, 1.0), oob_frac=0.0,
                 density=None):
        self.cdims = tuple(cdims)
        self.dims = tuple(dims)
        self.order = int(order)
        self.nb = int(nb) if nb is not None else (3 if order == 3 else 2)
        self.ns = ns
        self.ppc = ppc
        self.q = np.array(q[:ns], dtype=np.float64)
        self.m = np.array(m[:ns], dtype=np.float64)
        self.delh = tuple(delh)
        self.nchunk = int(np.prod(self.cdims))
        self.coord = np.asarray(coord, dtype=np.int32) if coord is not None else chunk_coords(self.cdims)
        self.M = tuple(d + 2 * self.nb for d in self.dims)
        self.seed = seed
        self.vth = vth
        self.amp = amp
        self.b0 = b0
        self.oob_frac = oob_frac
        self.density = density

    def ncell(self):
        return int(np.prod(self.dims))

    def field(self, k):
        """E/B of chunk k *including ghosts* drawn independently (call exchange_field to make the
        ghosts consistent)."""
        rng = np.random.default_rng([self.seed, 7, k])
        uf = rng.uniform(-self.amp, self.amp, size=self.M + (6,))
        uf[..., 5] += self.b0
        return uf

    def particles(self, k, s):
        rng = np.random.default_rng([self.seed, 1000 + 16 * k + s])
        n = self.ncell() * self.ppc
        if self.density is not None:
            n = int(n * self.density(self.coord[k], self.cdims))
        off = self.coord[k] * np.array(self.dims)
        xu = np.empty((n, 7), dtype=np.float64)
        ext = np.array(self.dims, dtype=np.float64) * np.array(self.delh)
        lo = off * np.array(self.delh)
        xu[:, 0] = lo[2] + rng.uniform(0, ext[2], n)
        xu[:, 1] = lo[1] + rng.uniform(0, ext[1], n)
        xu[:, 2] = lo[0] + rng.uniform(0, ext[0], n)
        # keep strictly inside [lo, hi)
        for c, a in ((0, 2), (1, 1), (2, 0)):
            xu[:, c] = np.minimum(xu[:, c], np.nextafter(lo[a] + ext[a], -np.inf))
        xu[:, 3:6] = rng.normal(0.0, self.vth[s], size=(n, 3))
        if self.oob_frac > 0:
            m = rng.uniform(size=n) < self.oob_frac
            xu[m, 0] += ext[2] * rng.choice([-1.0, 1.0], size=int(m.sum()))
        ids = (np.arange(n, dtype=np.int64) + (np.int64(k) << 32) + (np.int64(s) << 56))
        xu[:, 6] = ids.view(np.float64)
        return xu

    def total_particles(self):
        return self.nchunk * self.ns * self.ncell() * self.ppc
