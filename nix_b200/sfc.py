"""Chunk order along the generalised Hilbert ("gilbert") space-filling curve.

The reference orders chunks with sfc::get_map3d / get_map2d (sfc.cpp:51-169), an implementation of
J. Cerveny's published gilbert curve for arbitrary box sizes; ChunkMap exposes it as
id -> (cz, cy, cx) (chunkmap.cpp:176-191) and Balancer::assign_initial cuts the id range into
contiguous per-rank segments (balancer.cpp:101-124).  This module restates the published algorithm
(integer halving truncates towards zero, as in the reference's C++) so that Python callers -- bench.py,
the tests -- hand nixb200_domain_create the SAME table a nix application would.  Pinned against the
reference's own ChunkMap through tests/golden/sfc_coords.npz.
"""
import numpy as np


def _sgn(v):
    return (v > 0) - (v < 0)


def _half(v):
    return int(v / 2)  # truncation towards zero (C++ integer division)


def _gen2d(out, x, y, ax, ay, bx, by):
    w, h = abs(ax + ay), abs(bx + by)
    dax, day, dbx, dby = _sgn(ax), _sgn(ay), _sgn(bx), _sgn(by)
    if h == 1:
        for _ in range(w):
            out.append((x, y))
            x, y = x + dax, y + day
        return
    if w == 1:
        for _ in range(h):
            out.append((x, y))
            x, y = x + dbx, y + dby
        return
    ax2, ay2, bx2, by2 = _half(ax), _half(ay), _half(bx), _half(by)
    w2, h2 = abs(ax2 + ay2), abs(bx2 + by2)
    if 2 * w > 3 * h:
        if w2 % 2 and w > 2:
            ax2, ay2 = ax2 + dax, ay2 + day
        _gen2d(out, x, y, ax2, ay2, bx, by)
        _gen2d(out, x + ax2, y + ay2, ax - ax2, ay - ay2, bx, by)
    else:
        if h2 % 2 and h > 2:
            bx2, by2 = bx2 + dbx, by2 + dby
        _gen2d(out, x, y, bx2, by2, ax2, ay2)
        _gen2d(out, x + bx2, y + by2, ax, ay, bx - bx2, by - by2)
        _gen2d(out, x + (ax - dax) + (bx2 - dbx), y + (ay - day) + (by2 - dby), -bx2, -by2, -(ax - ax2), -(ay - ay2))


def _gen3d(out, p, a, b, c):
    x, y, z = p
    ax, ay, az = a
    bx, by, bz = b
    cx, cy, cz = c
    w, h, d = abs(ax + ay + az), abs(bx + by + bz), abs(cx + cy + cz)
    da = (_sgn(ax), _sgn(ay), _sgn(az))
    db = (_sgn(bx), _sgn(by), _sgn(bz))
    dc = (_sgn(cx), _sgn(cy), _sgn(cz))
    for n, dd, o1, o2 in ((w, da, h, d), (h, db, w, d), (d, dc, w, h)):
        if o1 == 1 and o2 == 1:
            for _ in range(n):
                out.append((x, y, z))
                x, y, z = x + dd[0], y + dd[1], z + dd[2]
            return
    a2 = [_half(v) for v in a]
    b2 = [_half(v) for v in b]
    c2 = [_half(v) for v in c]
    w2, h2, d2 = abs(sum(a2)), abs(sum(b2)), abs(sum(c2))
    if w2 % 2 and w > 2:
        a2 = [a2[i] + da[i] for i in range(3)]
    if h2 % 2 and h > 2:
        b2 = [b2[i] + db[i] for i in range(3)]
    if d2 % 2 and d > 2:
        c2 = [c2[i] + dc[i] for i in range(3)]

    def add(*vs):
        return tuple(sum(v[i] for v in vs) for i in range(3))

    def neg(v):
        return tuple(-t for t in v)

    def sub(u, v):
        return tuple(u[i] - v[i] for i in range(3))

    p = (x, y, z)
    a, b, c, a2, b2, c2 = tuple(a), tuple(b), tuple(c), tuple(a2), tuple(b2), tuple(c2)
    if 2 * w > 3 * h and 2 * w > 3 * d:  # wide: split along the major axis only
        _gen3d(out, p, a2, b, c)
        _gen3d(out, add(p, a2), sub(a, a2), b, c)
    elif 3 * h > 4 * d:  # do not split the third axis
        _gen3d(out, p, b2, c, a2)
        _gen3d(out, add(p, b2), a, sub(b, b2), c)
        _gen3d(out, add(p, sub(a, da), sub(b2, db)), neg(b2), c, neg(sub(a, a2)))
    elif 3 * d > 4 * h:  # do not split the second axis
        _gen3d(out, p, c2, a2, b)
        _gen3d(out, add(p, c2), a, b, sub(c, c2))
        _gen3d(out, add(p, sub(a, da), sub(c2, dc)), neg(c2), neg(sub(a, a2)), b)
    else:  # regular: split all three
        _gen3d(out, p, b2, c2, a2)
        _gen3d(out, add(p, b2), c, a2, sub(b, b2))
        _gen3d(out, add(p, sub(b2, db), sub(c, dc)), a, neg(b2), neg(sub(c, c2)))
        _gen3d(out, add(p, sub(a, da), b2, sub(c, dc)), neg(c), neg(sub(a, a2)), sub(b, b2))
        _gen3d(out, add(p, sub(a, da), sub(b2, db)), neg(b2), c2, neg(sub(a, a2)))


def _curve2d(n1, n0):
    """points (i0, i1) of the 2-D curve over a box with n0 columns (fast axis) and n1 rows"""
    out = []
    if n0 >= n1:
        _gen2d(out, 0, 0, n0, 0, 0, n1)
    else:
        _gen2d(out, 0, 0, 0, n1, n0, 0)
    return out


def chunk_coords(cdims):
    """id -> (cz, cy, cx) for a box of (Cz, Cy, Cx) chunks: nix::ChunkMap::get_coordinate
    (chunkmap.cpp:176-191) over sfc::get_map3d (sfc.cpp:97-169)."""
    cz, cy, cx = (int(v) for v in cdims)
    n = [cz, cy, cx]
    big = [i for i in range(3) if n[i] != 1]
    if len(big) == 3:
        pts = []
        if cx >= cy and cx >= cz:
            _gen3d(pts, (0, 0, 0), (cx, 0, 0), (0, cy, 0), (0, 0, cz))
        elif cy >= cx and cy >= cz:
            _gen3d(pts, (0, 0, 0), (0, cy, 0), (cx, 0, 0), (0, 0, cz))
        else:
            _gen3d(pts, (0, 0, 0), (0, 0, cz), (cx, 0, 0), (0, cy, 0))
        return np.array([(z, y, x) for (x, y, z) in pts], dtype=np.int32)
    if len(big) == 2:  # the reference treats the two non-trivial axes as (rows, columns) of get_map2d
        hi, lo = big  # hi = slower axis (z before y before x)
        out = np.zeros((cz * cy * cx, 3), dtype=np.int32)
        for i, (c0, c1) in enumerate(_curve2d(n[hi], n[lo])):
            out[i, lo], out[i, hi] = c0, c1
        return out
    out = np.zeros((cz * cy * cx, 3), dtype=np.int32)
    if len(big) == 1:
        out[:, big[0]] = np.arange(n[big[0]])
    return out


def rank_boundary(nchunk, nrank, load=None):
    """Contiguous curve segments per rank.  Without loads: what Balancer::assign_initial
    (balancer.cpp:101-124) yields for a uniform load vector (unittest/test_balancer.cpp:41-44:
    boundary[r] = r * nchunk / nrank)."""
    if load is None:
        return np.array([(nchunk * r) // nrank for r in range(nrank + 1)], dtype=np.int32)
    raise NotImplementedError("see nix_b200.balancer")
