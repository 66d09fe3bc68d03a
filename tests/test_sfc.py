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
