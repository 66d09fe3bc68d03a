"""nix_b200/balancer.py against the reference's own nix::Balancer (golden vectors written by
tests/golden/make_balancer_golden.py from balancer.cpp) and the reference's unit test."""
import os

import numpy as np

from nix_b200 import balancer

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "balancer.npz")


def test_assignments_equal_the_reference_balancer():
    g = np.load(GOLD)
    n = int(g["ncase"])
    assert n >= 20
    moved = 0
    for c in range(n):
        load, nrank = g[f"load_{c}"], int(g[f"nrank_{c}"])
        assert balancer.assign_initial(load, nrank) == g[f"initial_{c}"].tolist(), c
        b1 = balancer.assign(load, g[f"uniform_{c}"].tolist())
        assert b1 == g[f"step1_{c}"].tolist(), c
        assert balancer.assign(load, b1) == g[f"step2_{c}"].tolist(), c
        moved += b1 != g[f"uniform_{c}"].tolist()
    assert moved > n // 3


def test_uniform_load_gives_uniform_boundaries():
    # unittest/test_balancer.cpp:41-44
    assert balancer.assign_initial(np.ones(512), 8) == [64 * r for r in range(9)]
