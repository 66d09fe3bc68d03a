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


def _ascending(boundary, nchunk):
    """Balancer::is_boundary_ascending (balancer.cpp:166-180)"""
    return boundary[0] == 0 and boundary[-1] == nchunk and all(boundary[i + 1] > boundary[i] for i in range(1, len(boundary) - 1))


def _optimum(boundary, load):
    """Balancer::is_boundary_optimum (balancer.cpp:182-201): every boundary brackets its share of the cumulative load"""
    cum = np.concatenate([[0.0], np.cumsum(np.asarray(load, dtype=np.float64))])
    nr = len(boundary) - 1
    return all(cum[boundary[i]] <= i * cum[-1] / nr < cum[boundary[i] + 1] for i in range(1, nr))


def test_reference_assign_initial_properties():
    # unittest/test_balancer.cpp:24-62: 10 ranks x 20 chunks; homogeneous load -> boundary[i] == 20 i, ascending,
    # optimum; loads uniform in [0.5, 1.5) -> ascending and optimum
    nr, per = 10, 20
    b = balancer.assign_initial(np.ones(nr * per), nr)
    assert b == [per * i for i in range(nr + 1)] and _ascending(b, nr * per) and _optimum(b, np.ones(nr * per))
    rng = np.random.default_rng(2024)
    for _ in range(50):
        load = rng.uniform(0.5, 1.5, nr * per)
        b = balancer.assign_initial(load, nr)
        assert _ascending(b, nr * per) and _optimum(b, load)


def test_reference_assign_properties():
    # unittest/test_balancer.cpp:64-104: one sweep from the uniform boundaries; homogeneous load leaves them
    # where they are, a random load keeps them ascending
    nr, per = 10, 20
    uniform = [per * i for i in range(nr + 1)]
    assert balancer.assign(np.ones(nr * per), uniform) == uniform
    rng = np.random.default_rng(7)
    for _ in range(50):
        load = rng.uniform(0.5, 1.5, nr * per)
        assert _ascending(balancer.assign(load, uniform), nr * per)
