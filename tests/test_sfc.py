"""nix_b200/sfc.py against the reference's own ChunkMap (golden tables generated from the reference by
tests/golden/make_sfc_golden.py): the chunk order handed to nixb200_domain_create by bench.py and the
tests is the one a nix application would hand over."""
import os

import numpy as np

from nix_b200.sfc import chunk_coords, rank_boundary

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sfc_coords.npz")


def test_chunk_coords_equal_the_reference_chunkmap():
    g = np.load(GOLD)
    assert len(g.files) >= 25
    for name in g.files:
        cd = tuple(int(v) for v in name.split("_")[1:])
        assert np.array_equal(chunk_coords(cd), g[name].astype(np.int32)), cd


def test_curve_is_continuous_and_rank_segments_are_compact():
    """consecutive ids are face neighbours wherever the reference's curve is (even sizes), and the 8 equal
    segments of the 16^3 box are the eight 8^3 octants (what bench.py --gpus 8 partitions)"""
    for cd in [(8, 8, 8), (16, 16, 16), (4, 4, 2), (16, 8, 8)]:
        c = chunk_coords(cd)
        assert np.abs(np.diff(c, axis=0)).sum(axis=1).max() == 1
    c = chunk_coords((16, 16, 16))
    bd = rank_boundary(len(c), 8)
    for r in range(8):
        seg = c[bd[r]:bd[r + 1]]
        assert ((seg.max(axis=0) - seg.min(axis=0)) == 7).all()


def test_uniform_boundary_is_assign_initial():
    # unittest/test_balancer.cpp:41-44: boundary[i] == i * nchunk / nrank for equal loads
    assert rank_boundary(512, 8).tolist() == [0, 64, 128, 192, 256, 320, 384, 448, 512]
    assert rank_boundary(10, 4).tolist() == [0, 2, 5, 7, 10]


def _locality(coord, distmax2):
    """sfc::check_locality3d (sfc.cpp:206-227): consecutive ids at most sqrt(distmax2) apart"""
    d = np.diff(coord.astype(np.int64), axis=0)
    return bool(((d * d).sum(axis=1) <= distmax2).all())


def _index_is_a_permutation(coord, cd):
    """sfc::check_index (sfc.cpp:174-185): every cell of the box is visited exactly once"""
    flat = (coord[:, 0].astype(np.int64) * cd[1] + coord[:, 1]) * cd[2] + coord[:, 2]
    return bool(np.array_equal(np.sort(flat), np.arange(cd[0] * cd[1] * cd[2])))


def test_reference_sfc3d_properties_even():
    # unittest/test_sfc.cpp:58-70 "SFC3D even": sizes from {1, 4, 20, 100}, distance 1 between neighbours
    # (100 is run against one other size here: pure-Python recursion)
    sizes = [(z, y, x) for z in (1, 4, 20) for y in (1, 4, 20) for x in (1, 4, 20)] + [(1, 4, 100), (4, 100, 1), (100, 1, 20)]
    for cd in sizes:
        c = chunk_coords(cd)
        assert c.shape == (cd[0] * cd[1] * cd[2], 3)
        assert _locality(c, 1), cd
        assert _index_is_a_permutation(c, cd), cd


def test_reference_sfc3d_properties_odd():
    # unittest/test_sfc.cpp:71-112 "SFC3D odd-x / odd-y / odd-z": one odd size, distance^2 <= 2
    for odd in (3, 7, 9):
        for a in (4, 8, 16):
            for b in (4, 8, 16):
                for cd in ((a, b, odd), (a, odd, b), (odd, a, b)):  # (Cz, Cy, Cx)
                    c = chunk_coords(cd)
                    assert _locality(c, 2), cd
                    assert _index_is_a_permutation(c, cd), cd
