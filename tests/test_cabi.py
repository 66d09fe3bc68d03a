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


_NULL_CALLS = r"""
import ctypes as C, sys
sys.path.insert(0, sys.argv[1])
from nix_b200 import core
lib = core.load_library()
def default(t):
    if t in (C.c_int, C.c_int64): return 0
    if t is C.c_double: return 0.0
    return None
for name in core.SYMBOLS:
    f = getattr(lib, name)
    assert f.argtypes is not None or name in ("nixb200_last_error", "nixb200_version", "nixb200_launch_count"), name
    print("CALL", name, flush=True)
    r = f(*[default(t) for t in (f.argtypes or [])])
    print("RET", name, r if not isinstance(r, bytes) else 0, flush=True)
"""


def test_null_arguments_are_refused_not_dereferenced(tmp_path):
    """Error behaviour of the boundary (status codes, no exception, no crash crosses it): every entry
    point called with a NULL handle / NULL pointers returns non-zero (or -1 for the counting getters) and leaves a
    message; nothing dereferences the NULL.  Runs in a child process so that a fault would be reported, not fatal."""
    import subprocess
    import sys
    script = tmp_path / "null_calls.py"
    script.write_text(_NULL_CALLS)
    p = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=300)
    lines = p.stdout.strip().splitlines()
    assert p.returncode == 0, f"crashed in {lines[-1] if lines else '?'}: {p.stderr[-400:]}"
    ret = {ln.split()[1]: int(ln.split()[2]) for ln in lines if ln.startswith("RET")}
    assert set(ret) == set(core.SYMBOLS)
    ok_with_null = {"nixb200_last_error", "nixb200_version", "nixb200_launch_count",
                    "nixb200_domain_destroy", "nixb200_plan_destroy"}  # (destroying nothing is not an error)
    wrong = {n: r for n, r in ret.items() if n not in ok_with_null and r == 0}
    assert not wrong, wrong


def test_descriptor_is_validated_before_the_device():
    """A bad nixb200_domain_desc is reported as such on any machine (the checks are host logic and run before the
    device is looked for); a GOOD one still fails here, loudly, for want of a device (test_no_cpu_fallback)."""
    q, m = [-1.0, 1.0], [1.0, 25.0]
    bad = [
        (dict(order=4, nb=3), "order"),
        (dict(order=0, nb=2), "order"),
        (dict(order=3, nb=2), "boundary margin"),  # stencil of order 3 needs three ghost layers
        (dict(order=2, nb=1), "boundary margin"),
        (dict(order=2, nb=2, dims=(8, 1, 8)), "chunk dims"),
        (dict(order=2, nb=2, cdims=(2, 0, 2), id_range=(0, 1)), "cdims"),
        (dict(order=2, nb=2, delh=(1.0, 0.0, 1.0)), "cell sizes"),
        (dict(order=2, nb=2, cc=0.0), "speed of light"),
        (dict(order=2, nb=2, pusher=7), "pusher"),
        (dict(order=2, nb=2, id_range=(3, 3)), "empty domain"),
        (dict(order=2, nb=2, id_range=(0, 9)), "chunk id range"),
        (dict(order=2, nb=2, q=[], m=[]), "empty domain"),
    ]
    for kw, msg in bad:
        kw = dict(kw)
        args = dict(cdims=kw.pop("cdims", (2, 2, 2)), dims=kw.pop("dims", (8, 8, 8)), nb=kw.pop("nb"), order=kw.pop("order"),
                    q=kw.pop("q", q), m=kw.pop("m", m))
        with pytest.raises(core.NixB200Error, match=msg):
            core.Domain(args["cdims"], args["dims"], args["nb"], args["order"], args["q"], args["m"], **kw)


@pytest.mark.parametrize("dims,nb", [((8, 8, 8), 1), ((8, 8, 8), 2), ((16, 16, 16), 2), ((32, 32, 32), 3), ((6, 8, 10), 2)])
def test_halo_buffer_layout_equals_the_reference_mpibuffer(dims, nb):
    """a12 on the CPU: the library's MpiBuffer layout (27 slots in z, y, x order, running-sum addresses, empty centre
    slot) equals what the reference's Chunk::set_mpi_buffer (chunk.cpp:257-286) builds -- read from the oracle port
    and, where it is built, from the reference's own class -- for E/B (48 B per cell) and J (32 B per cell).
    SURVEY.md 8a quotes 666 624 / 444 416 bytes per exchange for 32^3 cells and two ghost layers."""
    from oracle import nixoracle as no
    for which in ("port", "ref"):
        if not no.available(which):
            continue
        c = no.Chunk(no.load(which), dims, nb, 2 if nb >= 2 else 1)
        for mode, omode in ((core.MODE_FIELD, no.MODE_FIELD), (core.MODE_CURRENT, no.MODE_CURRENT)):
            bs, ba = core.halo_layout_dims(dims, nb, mode)
            assert np.array_equal(bs, c.bufsize(omode)), (which, mode)
            assert np.array_equal(ba, c.bufaddr(omode)), (which, mode)
    bs, _ = core.halo_layout_dims((32, 32, 32), 2, core.MODE_FIELD)
    assert int(bs.sum()) == 666624
    bs, _ = core.halo_layout_dims((32, 32, 32), 2, core.MODE_CURRENT)
    assert int(bs.sum()) == 444416
    with pytest.raises(core.NixB200Error, match="field and current"):
        core.halo_layout_dims((8, 8, 8), 2, core.MODE_PARTICLE)
