"""The C++ host mirror (host/nixb200_host.hpp): GpuChunk : nix::Chunk, GpuInterface, GpuApplication
compiled against the REFERENCE's own headers, and host/_build/demo -- the reference's ChunkMap (Gilbert
space-filling curve), Chunk and XtensorParticle classes driving libnixb200.so.

CPU: the mirror compiles and links (here, where /root/reference exists) and the chunk order the demo
gets from nix::ChunkMap is a valid locality-preserving curve.  GPU: the demo's results -- obtained
through Chunk staging, i.e. what Chunk::pack / diagnostics see -- equal the CPU oracle bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from nix_b200.synth import Problem

from helpers import bits, oracle_domain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "host", "_build", "demo")


@pytest.fixture(scope="module")
def demo():
    if os.path.isdir("/root/reference"):
        from nix_b200 import build
        build.build()
        subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    if not os.path.exists(DEMO):
        pytest.skip("host/_build/demo not built (reference sources absent)")
    return DEMO


def chunkmap_coord(demo, cdims):
    out = subprocess.run([demo, "coord"] + [str(c) for c in cdims], check=True, capture_output=True, text=True).stdout
    return np.array([[int(v) for v in ln.split()] for ln in out.strip().splitlines()], dtype=np.int32)


def test_host_mirror_compiles_against_the_reference(demo):
    if os.path.isdir("/root/reference"):
        assert os.path.exists(os.path.join(ROOT, "host", "_build", "app_check.o"))  # GpuApplication : nix::Application


@pytest.mark.parametrize("cdims", [(2, 2, 2), (2, 4, 6), (8, 8, 8)])
def test_chunkmap_order_is_a_space_filling_curve(demo, cdims):
    """What the reference's own SFC tests assert (test_sfc.cpp:11-113): a permutation of the grid
    whose consecutive ids are face neighbours (distance^2 <= 1 for even sizes)."""
    coord = chunkmap_coord(demo, cdims)
    assert coord.shape == (int(np.prod(cdims)), 3)
    assert len({tuple(c) for c in coord}) == len(coord)
    assert np.all(coord >= 0) and np.all(coord < np.array(cdims))
    d2 = (np.diff(coord, axis=0) ** 2).sum(axis=1)
    assert d2.max() <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("order,cdims,n", [(2, (2, 2, 2), 8), (1, (2, 2, 4), 8), (3, (2, 2, 2), 8)])
def test_demo_equals_oracle(demo, oracle_port, gpu_lib, tmp_path, order, cdims, n):
    coord = chunkmap_coord(demo, cdims)
    prob = Problem(cdims, (n, n, n), order, ppc=8, seed=61 + order, vth=(0.35, 0.08), coord=coord)
    od = oracle_domain(oracle_port, prob, sort=True)
    # the demo's inputs: E/B with consistent ghosts is not required (make_domain exchanges), particles unsorted
    for k in range(prob.nchunk):
        prob.field(k).tofile(tmp_path / f"uf_{k}.bin")
        for s in range(prob.ns):
            prob.particles(k, s).tofile(tmp_path / f"xu_{k}_{s}.bin")
    steps = 3
    npmax = 4 * prob.ncell() * prob.ppc
    r = subprocess.run([demo, "run", str(tmp_path)] + [str(c) for c in cdims] +
                       [str(n), str(order), str(prob.nb), str(prob.ns), str(npmax), str(steps)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("ok"), r.stdout
    for _ in range(steps):
        od.step(0.5, 1.0)
    assert f"particles={od.total_particles()}" in r.stdout
    for k, c in enumerate(od.chunks):
        uj = np.fromfile(tmp_path / f"out_uj_{k}.bin").reshape(c.uj.shape)
        assert np.abs(uj - c.uj).max() <= 1e-12 * np.abs(c.uj).max()
        for s in range(prob.ns):
            xu = np.fromfile(tmp_path / f"out_xu_{k}_{s}.bin").reshape(-1, 7)
            ref = c.particles(s)
            assert xu.shape == ref.shape and np.array_equal(bits(xu), bits(ref)), f"chunk {k} species {s}"
