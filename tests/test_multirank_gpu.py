"""Multi-rank GPU parity: chunks partitioned over ranks along the id order, J / E/B halo and particle
migration across ranks over NCCL (csrc/peer.cu), against the single-process CPU oracle of the whole
box.  Strict mode => particles, counts and E/B bit-exact; J <= 1e-12 of its maximum.

The N-rank tests need N GPUs (`gpurun --gpus N`); on a 1-GPU box they skip and only the single-rank
pass through the multi-rank code path runs.  The logs of the 2/4/8-GPU runs of this file are kept
under profiles/ (r02_multirank_n*.log)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from nix_b200.synth import Problem  # noqa: E402

pytestmark = pytest.mark.gpu
JTOL = 1e-12


def _collect(q, procs, timeout):
    """One result per rank; a rank that dies or hangs (e.g. inside a collective its peer never
    entered) is reported and every process is reaped."""
    import queue
    res = []
    try:
        for _ in procs:
            res.append(q.get(timeout=timeout))
    except queue.Empty:
        got = {r[0] for r in res}
        res += [(i, "fail: no result (crashed or hung)") + (0,) * (len(res[0]) - 2 if res else 0)
                for i in range(len(procs)) if i not in got]
    finally:
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()
    return res


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _compare(rank, prob, od, gd, ids, what):
    from helpers import bits, ref_pcount_before_sort
    for k, i in enumerate(ids):
        c = od.chunks[i]
        uj = gd.get_current(k)
        scale = np.abs(c.uj).max() or 1.0
        assert np.abs(uj - c.uj).max() / scale < JTOL, f"{what} rank {rank} chunk {i}: J"
        assert np.array_equal(gd.get_field(k), c.uf), f"{what} rank {rank} chunk {i}: E/B"
        for s in range(prob.ns):
            ref, got = c.particles(s), gd.get_particles(k, s)
            assert got.shape == ref.shape, f"{what} rank {rank} chunk {i} sp {s}: Np {got.shape[0]} != {ref.shape[0]}"
            assert np.array_equal(bits(got), bits(ref)), f"{what} rank {rank} chunk {i} sp {s}: particles"
            assert np.array_equal(gd.get_pindex(k, s), c.pindex(s)), f"{what} rank {rank} chunk {i} sp {s}: pindex"
            assert np.array_equal(gd.get_pcount(k, s), ref_pcount_before_sort(c, s))


def _load(gd, prob, ids):
    for k, i in enumerate(ids):
        gd.set_field(k, prob.field(i))
    gd.exchange_field()
    for s in range(prob.ns):
        gd.set_particles(s, [prob.particles(i, s) for i in ids])
    gd.sort()


def test_single_rank_through_the_multirank_path(oracle_port, gpu_lib):
    """nrank = 1: no peers, but the step runs through peer_migrate (count tables, host sync)."""
    from nix_b200 import core
    from helpers import oracle_domain
    prob = Problem((2, 2, 2), (8, 8, 8), 2, ppc=8, seed=91, vth=(0.35, 0.08))
    od = oracle_domain(oracle_port, prob)
    gd = core.Domain(prob.cdims, prob.dims, prob.nb, prob.order, prob.q, prob.m, coord=prob.coord, strict_fp=True)
    gd.set_ranks([0, prob.nchunk], 0)
    _load(gd, prob, list(range(prob.nchunk)))
    for step in range(3):
        od.step(0.5, 1.0)
        gd.step(0.5)
        assert gd.check() == 0
        _compare(0, prob, od, gd, list(range(prob.nchunk)), f"step {step}")
    assert gd.peer_traffic() == dict(halo_cells_sent=0, particles_sent=0, particles_received=0)
    gd.close()


def _worker(rank, world, port, cdims, dims, order, steps, q):
    try:
        import torch
        import torch.distributed as dist
        from nix_b200 import core
        from oracle import nixoracle as no
        from helpers import oracle_domain
        torch.cuda.set_device(rank)
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        prob = Problem(cdims, dims, order, ppc=8, seed=93 + order, vth=(0.35, 0.08))
        bd = core.uniform_boundary(prob.nchunk, world)
        ids = list(range(int(bd[rank]), int(bd[rank + 1])))
        gd = core.Domain(prob.cdims, prob.dims, prob.nb, prob.order, prob.q, prob.m, coord=prob.coord,
                         id_range=(ids[0], ids[-1] + 1), device=rank, strict_fp=True)
        gd.set_ranks(bd, rank)
        gd.comm_init_torch()
        _load(gd, prob, ids)
        od = oracle_domain(no.load("port"), prob)  # whole box on the CPU of every rank
        _compare(rank, prob, od, gd, ids, "load")
        sent = 0
        for step in range(steps):
            od.step(0.5, 1.0)
            gd.step(0.5)
            assert gd.check() == 0, f"rank {rank}: device error bits"
            _compare(rank, prob, od, gd, ids, f"step {step}")
            tr = gd.peer_traffic()
            sent += tr["particles_sent"]
            assert tr["halo_cells_sent"] > 0
        tot = torch.tensor([gd.total_particles(), sent], dtype=torch.int64)
        dist.all_reduce(tot)
        assert int(tot[0]) == od.total_particles()
        assert int(tot[1]) > 0, "no particle crossed a rank boundary"
        gd.close()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: " + "".join(traceback.format_exception(exc))))


def _run_ranks(world, target, args, timeout=420):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=target, args=(r, world, port) + tuple(args) + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(q, procs, timeout)
    for rank, status in res:
        assert status == "ok", f"rank {rank}: {status}"


@pytest.mark.parametrize("cdims,dims,order", [((2, 2, 2), (8, 8, 8), 2), ((1, 2, 3), (6, 8, 10), 1),
                                              ((2, 2, 4), (8, 8, 8), 3), ((1, 1, 2), (8, 8, 8), 2)])
def test_two_ranks_equal_single_process_oracle(oracle_port, gpu_lib, cdims, dims, order):
    _run_ranks(2, _worker, (cdims, dims, order, 3))


# 4 and 8 ranks: contiguous segments of the reference's Gilbert curve (nix_b200/sfc.py) -- a rank's
# chunks touch several other ranks through faces, edges and corners, and some ranks are not neighbours
@pytest.mark.parametrize("cdims,dims,order", [((4, 4, 2), (8, 8, 8), 2), ((4, 4, 2), (8, 8, 8), 3),
                                              ((2, 3, 4), (6, 8, 10), 1)])
def test_four_ranks_equal_single_process_oracle(oracle_port, gpu_lib, cdims, dims, order):
    _run_ranks(4, _worker, (cdims, dims, order, 3))


@pytest.mark.parametrize("cdims,dims,order", [((4, 4, 2), (8, 8, 8), 2), ((4, 4, 2), (8, 8, 8), 3),
                                              ((4, 4, 4), (8, 8, 8), 2), ((2, 2, 2), (16, 16, 16), 2)])
def test_eight_ranks_equal_single_process_oracle(oracle_port, gpu_lib, cdims, dims, order):
    _run_ranks(8, _worker, (cdims, dims, order, 3))


def _growth_worker(rank, world, port, reserve, q):
    """Rank 0 owns a chunk full of particles that all fly in +x, rank 1 an empty one sized for nothing:
    rank 1's store must grow as they arrive (the reference resizes in pre_unpack,
    xtensor_halo3d.hpp:464-476)."""
    try:
        import torch
        import torch.distributed as dist
        from nix_b200 import core
        from oracle import nixoracle as no
        from helpers import oracle_domain
        torch.cuda.set_device(rank)
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        n = 30000
        prob = Problem((1, 1, 2), (8, 8, 8), 2, ppc=1, ns=1, seed=5, vth=(0.0,), amp=0.0, b0=0.0)

        def particles(k, s):
            m = n if k == 0 else 0
            rng = np.random.default_rng([77, k])
            xu = np.zeros((m, 7))
            xu[:, 0] = rng.uniform(0.0, 8.0, m) * (1 - 1e-12)
            xu[:, 1] = rng.uniform(0.0, 8.0, m) * (1 - 1e-12)
            xu[:, 2] = rng.uniform(0.0, 8.0, m) * (1 - 1e-12)
            xu[:, 3] = 50.0
            xu[:, 6] = np.arange(m, dtype=np.int64).view(np.float64)
            return xu
        prob.particles = particles
        bd = core.uniform_boundary(prob.nchunk, world)
        ids = [rank]
        gd = core.Domain(prob.cdims, prob.dims, prob.nb, prob.order, prob.q, prob.m, coord=prob.coord,
                         id_range=(rank, rank + 1), device=rank, strict_fp=True, capacity_factor=1.0)
        gd.set_ranks(bd, rank)
        gd.comm_init_torch()
        _load(gd, prob, ids)
        od = oracle_domain(no.load("port"), prob)
        cap0 = gd.capacity(0)
        if reserve:
            # 6 % of 30000 arrive per step: the first arrival has no history to warn the host
            gd.reserve(0, np_=4096)
        status = "ok"
        # without room reserved only ONE step is taken: a rank that refuses the next step would leave its
        # peer alone in the exchange (the caller's job is to check the error bits every step)
        for step in range(6 if reserve else 1):
            od.step(0.5, 1.0)
            gd.step(0.5)
            err = gd.check()
            if err:
                status = f"errbits {err}"
                continue
            _compare(rank, prob, od, gd, ids, f"growth step {step}")
        cap1 = gd.capacity(0)
        flag = torch.tensor([0 if status == "ok" else 1, int(cap1[0] > cap0[0])], dtype=torch.int64)
        gather = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gather, flag)
        gd.close()
        dist.barrier()
        dist.destroy_process_group()
        if reserve:
            assert status == "ok", status
            assert int(gather[1][1]) == 1, "rank 1's particle store did not grow"
        else:
            # rank 1 (sized for 1024 particles) receives 1800 at once: error bits, later steps refused,
            # no crash and no hang of the peer
            assert int(gather[1][0]) == 1, "the overflow went unnoticed"
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: " + "".join(traceback.format_exception(exc))))


@pytest.mark.parametrize("reserve", [True, False])
def test_two_ranks_particle_store_growth(oracle_port, gpu_lib, reserve):
    _run_ranks(2, _growth_worker, (reserve,))


def _rebalance_worker(rank, world, port, fp32, q):
    """Rank boundaries move twice during a run (chunks travel GPU to GPU in the reference's wire format,
    nixb200_domain_rebalance); after every step every rank still equals the single-process oracle."""
    try:
        import torch
        import torch.distributed as dist
        from nix_b200 import core
        from oracle import nixoracle as no
        from helpers import oracle_domain
        torch.cuda.set_device(rank)
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        prob = Problem((2, 2, 4), (8, 8, 8), 2, ppc=8, seed=123, vth=(0.35, 0.08),
                       density=lambda c, cd: 1.0 + 0.8 * np.sin(2 * np.pi * (c[2] + 0.5) / cd[2]))
        n = prob.nchunk
        plans = {2: [[0, 8, 16], [0, 5, 16], [0, 11, 16]],
                 4: [[0, 4, 8, 12, 16], [0, 3, 9, 11, 16], [0, 5, 8, 13, 16]]}[world]
        bd = plans[0]
        ids = list(range(bd[rank], bd[rank + 1]))
        gd = core.Domain(prob.cdims, prob.dims, prob.nb, prob.order, prob.q, prob.m, coord=prob.coord,
                         id_range=(ids[0], ids[-1] + 1), device=rank, strict_fp=not fp32, fp32=fp32)
        gd.set_ranks(bd, rank)
        gd.comm_init_torch()
        _load(gd, prob, ids)
        od = oracle_domain(no.load("port"), prob)
        step = 0
        for phase, bd in enumerate(plans):
            if phase > 0:
                gd.rebalance(bd, rank)
                ids = list(range(bd[rank], bd[rank + 1]))
                assert gd.nchunk == len(ids)
                if not fp32:
                    _compare(rank, prob, od, gd, ids, f"after rebalance {phase}")
            for _ in range(2):
                od.step(0.5, 1.0)
                gd.step(0.5)
                assert gd.check() == 0, f"rank {rank}: device error bits"
                if not fp32:
                    _compare(rank, prob, od, gd, ids, f"phase {phase} step {step}")
                step += 1
        tot = torch.tensor([gd.total_particles()], dtype=torch.int64)
        dist.all_reduce(tot)
        assert int(tot[0]) == od.total_particles()
        if fp32:
            for k, i in enumerate(ids):
                c = od.chunks[i]
                assert np.abs(gd.get_current(k) - c.uj).max() / np.abs(c.uj).max() < 1e-2
                for s in range(prob.ns):
                    assert abs(len(gd.get_particles(k, s)) - len(c.particles(s))) <= 2
        gd.close()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: " + "".join(traceback.format_exception(exc))))


@pytest.mark.parametrize("world,fp32", [(2, False), (4, False), (2, True)])
def test_device_to_device_rebalance(oracle_port, gpu_lib, world, fp32):
    _run_ranks(world, _rebalance_worker, (fp32,))


def test_single_rank_rebalance_and_wire_record(oracle_port, gpu_lib):
    """nrank = 1: the rebuild path alone (nothing moves) keeps the state bit for bit; the wire payload has the
    size the reference's XtensorParticle::pack has for the same particle counts (the content is checked through
    the reference's own unpack in tests/test_host_cpp.py::test_demo_equals_oracle)."""
    from nix_b200 import core
    from helpers import oracle_domain
    prob = Problem((2, 2, 2), (8, 8, 8), 2, ppc=8, seed=5, vth=(0.35, 0.08))
    od = oracle_domain(oracle_port, prob)
    gd = core.Domain(prob.cdims, prob.dims, prob.nb, prob.order, prob.q, prob.m, coord=prob.coord, strict_fp=True)
    gd.set_ranks([0, prob.nchunk], 0)
    _load(gd, prob, list(range(prob.nchunk)))
    for step in range(2):
        od.step(0.5, 1.0)
        gd.step(0.5)
    gd.rebalance([0, prob.nchunk], 0)
    _compare(0, prob, od, gd, list(range(prob.nchunk)), "after a no-op rebalance")
    od.step(0.5, 1.0)
    gd.step(0.5)
    _compare(0, prob, od, gd, list(range(prob.nchunk)), "step after it")
    cells = int(np.prod(gd.M))
    for k, c in enumerate(od.chunks):
        w = gd.wire_pack(k)
        want = 8 + cells * 10 * 8
        for s in range(prob.ns):
            npt = ((c.np(s) + 128) // 128) * 128
            want += 175 + npt * 7 * 8 * 2 + npt * 4 + (cells + 1) * 4 + (cells + 1) * 8 * 4
        assert len(w) == want
        assert np.frombuffer(w[:8].tobytes(), dtype=np.int32).tolist() == [prob.order, prob.ns]
        uf = np.frombuffer(w[8:8 + cells * 48].tobytes(), dtype=np.float64).reshape(c.uf.shape)
        assert np.array_equal(uf, c.uf)
        off = 8 + cells * 80
        for s in range(prob.ns):
            hdr = w[off:off + 175].tobytes()
            npt, npp, ng = np.frombuffer(hdr[:12], dtype=np.int32)
            assert npp == c.np(s) and ng == cells
            xu = np.frombuffer(w[off + 175:off + 175 + npt * 56].tobytes(), dtype=np.float64).reshape(npt, 7)
            assert np.array_equal(xu[:npp].view(np.int64), c.particles(s).view(np.int64))
            pidx = np.frombuffer(w[off + 175 + npt * 116:off + 175 + npt * 116 + (cells + 1) * 4].tobytes(), dtype=np.int32)
            assert np.array_equal(pidx, c.pindex(s))
            off += 175 + npt * 116 + (cells + 1) * 36
    gd.close()


def test_two_domains_on_two_devices_in_one_process(oracle_port, gpu_lib):
    """ADVICE r01 (push_deposit.cu:1172): kernel attributes are per device and every entry point must run on the
    domain's device whatever the caller's current device is.  Two independent domains on cuda:0 and cuda:1,
    stepped alternately from one thread while the current device points at the OTHER one."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from nix_b200 import core
    from helpers import assert_particles_equal, oracle_domain
    probs = [Problem((2, 2, 2), (8, 8, 8), 2, ppc=8, seed=11, vth=(0.3, 0.05)),
             Problem((1, 2, 2), (8, 8, 8), 3, ppc=6, seed=12, vth=(0.3, 0.05))]
    ods = [oracle_domain(oracle_port, p) for p in probs]
    gds = []
    for dev, p in enumerate(probs):
        torch.cuda.set_device(1 - dev)  # deliberately the wrong current device
        gd = core.Domain(p.cdims, p.dims, p.nb, p.order, p.q, p.m, coord=p.coord, device=dev, strict_fp=True)
        _load(gd, p, list(range(p.nchunk)))
        gds.append(gd)
        assert torch.cuda.current_device() == 1 - dev, "the library must restore the caller's device"
    for step in range(3):
        for dev in (0, 1):
            torch.cuda.set_device(1 - dev)
            ods[dev].step(0.5, 1.0)
            gds[dev].step(0.5)
            assert gds[dev].check() == 0
    for dev in (0, 1):
        torch.cuda.set_device(1 - dev)
        assert_particles_equal(ods[dev], gds[dev], f"device {dev}")
        gds[dev].close()
