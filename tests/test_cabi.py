"""The drop-in boundary without a GPU: libnixb200.so loads, exports every entry point include/nixb200.h declares
(and the ctypes mirror binds exactly that set), the host-only entry points work, and the product path fails
LOUDLY on a machine without an sm_100 device -- there is no CPU fallback to fall back to."""
import ctypes
import os
import re

import numpy as np
import pytest

from nix_b200 import core

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    hdr = open(os.path.join(ROOT, "include", "nixb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return set(re.findall(r"\b(nixb200_[a-z0-9_]+)\s*\(", hdr))


def test_library_exports_every_declared_entry_point():
    names = header_functions()
    assert len(names) >= 60
    lib = ctypes.CDLL(core.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, f"declared in include/nixb200.h but not exported: {missing}"
    assert names == set(core.SYMBOLS), (sorted(names - set(core.SYMBOLS)), sorted(set(core.SYMBOLS) - names))


def test_every_entry_point_cites_the_reference():
    """the header documents, next to its entry points, the reference interface each one stands behind"""
    hdr = open(os.path.join(ROOT, "include", "nixb200.h")).read()
    for cite in ("application.hpp:343-346", "chunk.hpp:435-586", "xtensor_particle.hpp:260-357", "xtensor_halo3d.hpp:251-557",
                 "chunkmap.cpp:156-164", "balancer.hpp:122-332", "xtensor_packer3d.hpp:62-82", "primitives.hpp:165-189",
                 "chunk.cpp:257-286", "xtensor_particle.hpp:128-169"):
        assert cite in hdr, cite


def test_host_only_entry_points():
    lib = core.load_library()
    assert lib.nixb200_version().decode().startswith("nixb200")
    assert core.launch_count() >= 0
    rc, mv = core.rebalance_moves((8, 16), (5, 16))
    assert rc == 0 and mv == [(8, 8), (16, 16), (5, 8), (16, 16), (8, 16)]
    rc, _ = core.rebalance_moves((0, 4), (6, 9))
    assert rc != 0  # the old and the new range must overlap
    plan = core.Plan((2, 2, 2), (8, 8, 8), 2, core.chunk_coords((2, 2, 2)), [0, 4, 8], 0)
    assert [p["rank"] for p in plan.peers] == [1]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: nothing to refuse")
    with pytest.raises(core.NixB200Error, match="CUDA|device"):
        core.Domain((1, 1, 1), (8, 8, 8), 2, 2, [-1.0], [1.0])
    x = np.zeros(4)
    with pytest.raises(core.NixB200Error):
        core.shape_eval(0, 2, x, x, 1.0)
