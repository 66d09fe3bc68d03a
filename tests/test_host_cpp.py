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


@pytest.mark.parametrize("dims,nb,counts", [((8, 8, 8), 2, [100, 7]), ((16, 16, 16), 2, [262144, 262144]),
                                            ((32, 32, 32), 3, [0, 1, 127, 128, 129]), ((6, 8, 10), 2, [4095])])
def test_wire_payload_size_equals_the_reference_pack(demo, dims, nb, counts):
    """Size of the chunk payload the DEVICE assembles (nixb200_chunk_wire_pack; layout arithmetic =
    nixb200_wire_size_dims, host logic) against the reference's OWN pack() in query mode on containers of exactly
    that many particles (`demo wiresize`: nix::Chunk::pack + XtensorParticle::pack, chunk.cpp:18-60,
    xtensor_particle.hpp:128-169).  The CONTENT is compared on the GPU (test_demo_equals_oracle: the reference's
    unpack reads the device-made record)."""
    from nix_b200 import core
    out = subprocess.run([demo, "wiresize"] + [str(v) for v in dims] + [str(nb), "2"] + [str(n) for n in counts],
                         check=True, capture_output=True, text=True).stdout.split()
    header, payload = int(out[0]), int(out[1])
    assert header > 0
    assert core.wire_size_dims(dims, nb, counts) == payload
    # and the pieces, spelled out: order, ns, uf, uj, then per species 175 bytes of scalars, xu + xv [Np_total][7],
    # gindex [Np_total], pindex [Ng + 1], pcount [Ng + 1][8] with Np_total = ((Np + 128) / 128) * 128
    cells = int(np.prod([d + 2 * nb for d in dims]))
    want = 8 + cells * 80
    for n in counts:
        npt = (n + 128) // 128 * 128
        want += 175 + npt * (2 * 56 + 4) + (cells + 1) * 36
    assert payload == want


@pytest.mark.parametrize("dims,nb,delh,offset,gdims,q,np_", [
    ((8, 8, 8), 2, (1.0, 1.0, 1.0), (0, 0, 0), (16, 16, 16), -1.0, 0),
    ((8, 6, 10), 2, (0.5, 1.25, 2.0), (16, 6, 30), (32, 24, 40), -1.0, 100),
    ((32, 32, 32), 3, (0.1, 0.1, 0.1), (224, 96, 0), (256, 256, 256), 1.0, 524288),
    ((16, 16, 16), 2, (1.0, 1.0, 1.0), (112, 0, 48), (128, 128, 128), 1.0, 127)])
def test_wire_particle_header_equals_the_reference_pack(demo, dims, nb, delh, offset, gdims, q, np_):
    """The 175 scalar bytes that open a species in the device-made chunk record (made on the host side of the
    library: Np_total, Np, Ng, q, m, the three has_dim flags, Lb / Ub per axis, cell sizes, the chunk's and the box's
    coordinate ranges) against the first 175 bytes of the reference's OWN XtensorParticle::pack
    (xtensor_particle.hpp:130-159) for a chunk of the same geometry (`demo wirehdr`): byte for byte, including
    anisotropic cells and a chunk away from the origin."""
    from nix_b200 import core
    args = [str(v) for v in dims] + [str(nb)] + [repr(float(v)) for v in delh] + [str(v) for v in offset] + \
        [str(v) for v in gdims] + [repr(float(q)), str(np_)]
    ref = subprocess.run([demo, "wirehdr"] + args, check=True, capture_output=True, text=True).stdout.strip()
    assert len(ref) == 350
    assert core.wire_particle_header(dims, nb, delh, offset, gdims, q, 25.0, np_).hex() == ref


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


APP = os.path.join(ROOT, "host", "_build", "app_main")


def _write_app_inputs(tmp_path, prob, extra=None, interval=2):
    import json
    for k in range(prob.nchunk):
        prob.field(k).tofile(tmp_path / f"uf_{k}.bin")
        for s in range(prob.ns):
            prob.particles(k, s).tofile(tmp_path / f"xu_{k}_{s}.bin")
    opt = {"input": str(tmp_path), "order": prob.order, "nb": prob.nb, "strict": True,
           "species": [[float(q), float(m)] for q, m in zip(prob.q, prob.m)], "np_max": 4 * prob.ncell() * prob.ppc}
    opt.update(extra or {})
    cfg = {
        "application": {"basedir": str(tmp_path), "log": {"prefix": "log", "path": ".", "interval": 100},
                        "rebalance": {"loglevel": 1, "interval": interval}, "mpistream": False, "option": opt},
        "diagnostic": [],
        "parameter": {"Nx": prob.cdims[2] * prob.dims[2], "Ny": prob.cdims[1] * prob.dims[1],
                      "Nz": prob.cdims[0] * prob.dims[0], "Cx": prob.cdims[2], "Cy": prob.cdims[1],
                      "Cz": prob.cdims[0], "delt": 0.5, "delh": 1.0},
    }
    (tmp_path / "config.json").write_text(json.dumps(cfg, indent=1))


def test_application_binary_is_built(demo):
    """GpuApplication is no longer compile-checked only: host/_build/app_main links the reference's own
    application.cpp / balancer.cpp / chunk.cpp / chunkmap.cpp / sfc.cpp / nixio.cpp with FileApplication :
    GpuApplication (and the single-process MPI stand-in where MPI is absent)."""
    if os.path.isdir("/root/reference"):
        assert os.path.exists(APP)


@pytest.mark.gpu
@pytest.mark.parametrize("order,field_solver", [(2, False), (3, False), (2, True)])
def test_application_main_equals_oracle(demo, oracle_port, gpu_lib, tmp_path, order, field_solver):
    """nix::Application::main() of the reference -- initialize, setup_chunks (Chunk::setup through the
    factory), then diagnostic / push / rebalance / take_log / increment_time per step, finalize -- with
    GpuApplication::push() on the device, over 6 steps = 3 rebalance intervals.  The final particles, as the
    Chunk staging (what pack(), checkpoints and diagnostics read) sees them, equal the oracle's bit for bit."""
    import json
    if not os.path.exists(APP):
        pytest.skip("host/_build/app_main not built (reference sources absent)")
    cdims, n = (2, 2, 2), 8
    coord = chunkmap_coord(demo, cdims)
    prob = Problem(cdims, (n, n, n), order, ppc=8, seed=71 + order, vth=(0.35, 0.08), coord=coord)
    cfj = 0.4
    _write_app_inputs(tmp_path, prob, extra={"field_solver": field_solver, "cfj": cfj})
    od = oracle_domain(oracle_port, prob, sort=True)
    # is_push_needed: curtime < tmax + delt (application.cpp:427-433) -> tmax = 2.5 with delt = 0.5 gives 6 steps
    r = subprocess.run([APP, "-c", str(tmp_path / "config.json"), "--tmax", "2.5"], capture_output=True, text=True,
                       timeout=300, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    summary = json.loads((tmp_path / "app_summary.json").read_text())
    assert summary["steps"] == 6 and summary["curstep"] == 6
    assert summary["rebalance_calls_that_ran"] >= 2      # steps 2 and 4 (curstep > 0 and curstep % interval == 0)
    assert summary["domain_builds"] == 1                 # a one-rank rebalance moves nothing: no rebuild, no download
    assert summary["launches"] > 0
    for _ in range(6):
        if field_solver:
            od.step_em(0.5, 1.0, cfj)
        else:
            od.step(0.5, 1.0)
    assert summary["particles"] == od.total_particles()
    for k, c in enumerate(od.chunks):
        xus = [np.fromfile(tmp_path / f"out_xu_{k}_{s}.bin").reshape(-1, 7) for s in range(prob.ns)]
        if not field_solver:
            for s in range(prob.ns):
                ref = c.particles(s)
                assert xus[s].shape == ref.shape and np.array_equal(bits(xus[s]), bits(ref)), f"chunk {k} species {s}"
            assert np.array_equal(np.fromfile(tmp_path / f"out_uf_{k}.bin").reshape(c.uf.shape), c.uf)
        else:
            # the deposit sums in another order (1e-12 of max J), E/B and then the particles inherit it
            uf = np.fromfile(tmp_path / f"out_uf_{k}.bin").reshape(c.uf.shape)
            assert np.abs(uf - c.uf).max() <= 1e-11 * np.abs(c.uf).max()
            for s in range(prob.ns):
                ref = c.particles(s)
                assert xus[s].shape == ref.shape
                o1 = np.argsort(np.ascontiguousarray(xus[s][:, 6]).view(np.int64))
                o2 = np.argsort(np.ascontiguousarray(ref[:, 6]).view(np.int64))
                assert np.abs(xus[s][o1, :6] - ref[o2, :6]).max() < 1e-10
        uj = np.fromfile(tmp_path / f"out_uj_{k}.bin").reshape(c.uj.shape)
        assert np.abs(uj - c.uj).max() <= 1e-11 * np.abs(c.uj).max()
