"""Host logic of the multi-rank path on CPU (no GPU): the plan of csrc/peer.cu (nix_b200.core.Plan) and
the rank partition, exercised with world_size-2/3 `gloo` process groups.  Data plane = the CPU
oracle split over ranks (tests/multirank_oracle.py); the result must equal the single-process
oracle domain bit for bit, for every chunk a rank owns."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from nix_b200 import core  # noqa: E402
from nix_b200.synth import Problem  # noqa: E402


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


def test_uniform_boundary_matches_balancer_known_answer():
    # unittest/test_balancer.cpp:41-44: uniform load over 8 ranks, 160 chunks -> boundary[i] == 20 i
    assert list(core.uniform_boundary(160, 8)) == [20 * i for i in range(9)]


@pytest.mark.parametrize("cdims,nrank", [((2, 2, 4), 2), ((3, 2, 2), 3), ((4, 4, 4), 8), ((2, 1, 2), 2), ((1, 1, 2), 2),
                                         ((2, 2, 4), [0, 4, 9, 14, 16])])
def test_plan_pairs_up(cdims, nrank):
    """What rank r sends to rank p is, entry by entry, what p expects from r; every slab whose
    neighbour lives elsewhere is listed exactly once.  The last case is the uneven boundary array of the
    reference's own ChunkMap tests (unittest/test_chunkmap.cpp:127-160: Cz, Cy, Cx = 2, 2, 4, boundary
    {0, 4, 9, 14, 16}; an id belongs to the rank whose range holds it, chunkmap.cpp:156-164)."""
    prob = Problem(cdims, (8, 6, 4), 2, ppc=1)
    nchunk = int(np.prod(cdims))
    if isinstance(nrank, list):
        bd, nrank = np.array(nrank), len(nrank) - 1
        assert [int(np.searchsorted(bd, i, side="right") - 1) for i in (0, 3, 4, 8, 9, 13, 14, 15)] == [0, 0, 1, 1, 2, 2, 3, 3]
    else:
        bd = core.uniform_boundary(nchunk, nrank)
    plans = [core.Plan(cdims, prob.dims, prob.nb, prob.coord, bd, r) for r in range(nrank)]
    owner = lambda i: int(np.searchsorted(bd, i, side="right") - 1)  # noqa: E731  ChunkMap::get_rank
    grid2id = {tuple(int(v) for v in c): i for i, c in enumerate(prob.coord)}
    for r, pl in enumerate(plans):
        ranks = [p["rank"] for p in pl.peers]
        assert ranks == sorted(set(ranks)) and r not in ranks
        expect = set()
        for i in range(bd[r], bd[r + 1]):
            for d in range(27):
                if d == 13:
                    continue
                e = (d // 9 - 1, (d // 3) % 3 - 1, d % 3 - 1)
                nid = grid2id[tuple((int(prob.coord[i][a]) + e[a]) % cdims[a] for a in range(3))]
                if owner(nid) != r:
                    expect.add((owner(nid), i, d))
        got = {(p["rank"], i, d) for p in pl.peers for (i, d, _) in p["send"]}
        assert got == expect
        assert {(p["rank"], i, d) for p in pl.peers for (i, d, _) in p["recv"]} == expect
        for p in pl.peers:
            back = next(q for q in plans[p["rank"]].peers if q["rank"] == r)
            assert len(p["send"]) == len(back["recv"])
            for (i, d, n), (j, e, m) in zip(p["send"], back["recv"]):
                assert e == 26 - d and n == m
                dd = (d // 9 - 1, (d // 3) % 3 - 1, d % 3 - 1)
                assert grid2id[tuple((int(prob.coord[i][a]) + dd[a]) % cdims[a] for a in range(3))] == j


def test_plan_rejects_bad_boundary():
    prob = Problem((2, 2, 2), (8, 8, 8), 2, ppc=1)
    with pytest.raises(core.NixB200Error):
        core.Plan(prob.cdims, prob.dims, prob.nb, prob.coord, [0, 4, 7], 0)
    with pytest.raises(core.NixB200Error):
        core.Plan(prob.cdims, prob.dims, prob.nb, prob.coord, [0, 5, 3, 8], 1)


def _worker(rank, world, port, cdims, dims, order, steps, q):
    try:
        import torch.distributed as dist
        from oracle import nixoracle as no
        from helpers import bits, oracle_domain
        from multirank_oracle import RankOracle
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        lib = no.load("port")
        prob = Problem(cdims, dims, order, ppc=6, seed=77, vth=(0.35, 0.08))
        bd = core.uniform_boundary(prob.nchunk, world)
        ro = RankOracle(lib, prob, bd, rank)
        ro.load()
        ro.exchange(no.MODE_FIELD)
        ro.sort_only()
        full = oracle_domain(lib, prob)  # the whole box in this process: the reference result
        moved = 0
        for _ in range(steps):
            ro.step(0.5, 1.0)
            full.step(0.5, 1.0)
        for i in ro.ids:
            a, b = ro.chunks[i], full.chunks[i]
            assert np.array_equal(a.uf, b.uf), f"rank {rank} chunk {i}: E/B differ"
            assert np.array_equal(bits(a.uj), bits(b.uj)), f"rank {rank} chunk {i}: J differs"
            for s in range(prob.ns):
                pa, pb = a.particles(s), b.particles(s)
                assert pa.shape == pb.shape and np.array_equal(bits(pa), bits(pb)), f"rank {rank} chunk {i} sp {s}"
                moved += int((pa[:, 6].view(np.int64) >> 32 & 0xFFFFFF != i).sum())
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", moved))
    except Exception as exc:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: " + "".join(traceback.format_exception(exc)), 0))


@pytest.mark.parametrize("world,cdims,dims,order", [(2, (2, 2, 2), (8, 8, 8), 2), (2, (1, 2, 3), (6, 8, 10), 1),
                                                    (3, (3, 2, 2), (8, 8, 8), 3)])
def test_rank_split_oracle_equals_single_process(oracle_port, world, cdims, dims, order):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cdims, dims, order, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(q, procs, 240)
    for rank, status, _ in res:
        assert status == "ok", f"rank {rank}: {status}"
    assert sum(m for _, _, m in res) > 0, "no particle changed chunk: the migration path was not exercised"


def test_rebalance_moves_are_consistent_between_ranks():
    """Host logic of nixb200_domain_rebalance (no device): for every boundary move the reference's Balancer makes
    on the golden load vectors, what rank r sends to r+1 is exactly what r+1 receives from r (and vice versa),
    every chunk has exactly one owner afterwards, and old and new range of a rank overlap (chunks move to
    rank-1 / rank+1 only, balancer.hpp:122-332)."""
    import numpy as np
    from nix_b200 import balancer, core
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "balancer.npz"))
    checked = 0
    for c in range(int(g["ncase"])):
        load = g[f"load_{c}"]
        old = g[f"uniform_{c}"].tolist()
        for _ in range(3):
            new = balancer.assign(load, old)
            nr = len(old) - 1
            mv, rcs = [], []
            for r in range(nr):
                rc, m = core.rebalance_moves((old[r], old[r + 1]), (new[r], new[r + 1]))
                rcs.append(rc)
                mv.append(m)
            if any(new[r + 1] <= new[r] for r in range(nr)):
                # one sweep of the reference's algorithm can leave a rank without chunks when the load is a spike
                # (balancer.cpp:26-63 bounds every boundary by its OLD neighbours only); nixb200_domain_rebalance
                # refuses such a range instead of building an empty domain
                assert any(rcs)
                break
            assert not any(rcs), (c, old, new)
            owned = np.zeros(len(load), dtype=int)
            for r in range(nr):
                sl, sr, rl, rr, keep = mv[r]
                if r + 1 < nr:
                    assert sr == mv[r + 1][2] or (sr[1] <= sr[0] and mv[r + 1][2][1] <= mv[r + 1][2][0])
                    assert mv[r + 1][0] == rr or (mv[r + 1][0][1] <= mv[r + 1][0][0] and rr[1] <= rr[0])
                else:
                    assert sr[1] <= sr[0] and rr[1] <= rr[0]
                if r == 0:
                    assert sl[1] <= sl[0] and rl[1] <= rl[0]
                for a, b in (rl, keep, rr):
                    if b > a:
                        owned[a:b] += 1
                assert (min(x for x, y in (rl, keep, rr) if y > x), max(y for x, y in (rl, keep, rr) if y > x)) == (new[r], new[r + 1])
            assert (owned == 1).all()
            checked += old != new
            old = new
    assert checked > 20


@pytest.mark.parametrize("world", [2, 4, 8])
def test_plan_of_the_benchmarked_boxes(world):
    """The boxes `bench.py --gpus N` runs (cfg2: 8x8x8 chunks of 16^3 per GPU, repeated z, y, x; ids along the
    reference's Gilbert curve, one contiguous segment per rank): every rank's plan pairs up with its peers'
    entry by entry, every slab that crosses a rank boundary appears exactly once, and the neighbour relation is
    symmetric (host logic only -- the N = 4 box is not covered by any GPU test of its own)."""
    saved = os.dup(1)  # (bench.py points file descriptor 1 at stderr when it is loaded: its stdout carries ONE line)
    try:
        import bench
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    gcd, coord = bench.global_box((8, 8, 8), world)
    assert int(np.prod(gcd)) == 512 * world
    nchunk = 512 * world
    bd = core.uniform_boundary(nchunk, world)
    assert list(bd) == [512 * r for r in range(world + 1)]
    dims, nb = (16, 16, 16), 2
    plans = [core.Plan(gcd, dims, nb, coord, bd, r) for r in range(world)]
    grid2id = {tuple(int(v) for v in c): i for i, c in enumerate(coord)}
    assert len(grid2id) == nchunk  # the curve visits every chunk once
    owner = np.searchsorted(bd, np.arange(nchunk), side="right") - 1
    for r, pl in enumerate(plans):
        ranks = [p["rank"] for p in pl.peers]
        assert ranks == sorted(set(ranks)) and r not in ranks
        expect = set()
        for i in range(bd[r], bd[r + 1]):
            c = coord[i]
            for d in range(27):
                if d == 13:
                    continue
                nid = grid2id[((int(c[0]) + d // 9 - 1) % gcd[0], (int(c[1]) + (d // 3) % 3 - 1) % gcd[1],
                               (int(c[2]) + d % 3 - 1) % gcd[2])]
                if owner[nid] != r:
                    expect.add((int(owner[nid]), i, d))
        assert {(p["rank"], i, d) for p in pl.peers for (i, d, _) in p["send"]} == expect
        for p in pl.peers:
            back = next(q for q in plans[p["rank"]].peers if q["rank"] == r)
            assert len(p["send"]) == len(back["recv"]) and len(p["recv"]) == len(back["send"])
            for (i, d, n), (j, e, m) in zip(p["send"], back["recv"]):
                assert e == 26 - d and n == m
                assert grid2id[((int(coord[i][0]) + d // 9 - 1) % gcd[0], (int(coord[i][1]) + (d // 3) % 3 - 1) % gcd[1],
                                (int(coord[i][2]) + d % 3 - 1) % gcd[2])] == j


def _rebalance_worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from oracle import nixoracle as no
        from helpers import bits, oracle_domain
        from multirank_oracle import RankOracle
        from nix_b200 import balancer
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        lib = no.load("port")
        # density rising along x: the uniform boundaries are not the balanced ones
        prob = Problem((1, 2, 6), (8, 8, 8), 2, ppc=4, seed=55, vth=(0.35, 0.08),
                       density=lambda c, cd: 0.4 + 1.6 * c[2] / (cd[2] - 1))
        bd = [int(v) for v in core.uniform_boundary(prob.nchunk, world)]
        ro = RankOracle(lib, prob, bd, rank)
        ro.load()
        ro.exchange(no.MODE_FIELD)
        ro.sort_only()
        full = oracle_domain(lib, prob)
        history = [list(bd)]
        for rnd in range(3):
            for _ in range(2):
                ro.step(0.5, 1.0)
                full.step(0.5, 1.0)
            # Chunk::load = particles per chunk (what GpuApplication::push() reports), all-gathered like
            # Balancer::assign's input (balancer.hpp:116)
            load = np.array([sum(full.chunks[i].np(s) for s in range(prob.ns)) for i in range(prob.nchunk)], dtype=np.float64)
            new = balancer.assign(load, bd)
            if new != bd:
                ro.rebalance(new)
                bd = new
                history.append(list(bd))
        for i in ro.ids:
            a, b = ro.chunks[i], full.chunks[i]
            assert np.array_equal(a.uf, b.uf), f"rank {rank} chunk {i}: E/B differ"
            assert np.array_equal(bits(a.uj), bits(b.uj)), f"rank {rank} chunk {i}: J differs"
            for s in range(prob.ns):
                pa, pb = a.particles(s), b.particles(s)
                assert pa.shape == pb.shape and np.array_equal(bits(pa), bits(pb)), f"rank {rank} chunk {i} sp {s}"
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", len(history) - 1))
    except Exception as exc:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: " + "".join(traceback.format_exception(exc)), 0))


def test_rank_split_oracle_with_rebalancing_equals_single_process(oracle_port):
    """world_size-3 `gloo` run of the REBALANCE path's host logic: Balancer::assign (nix_b200/balancer.py) on the
    per-chunk particle counts of a non-uniform box moves the rank boundaries between steps, chunks travel to
    rank-1 / rank+1 as nixb200_rebalance_moves says, the plan is rebuilt -- and every rank still equals the
    single-process oracle bit for bit afterwards."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    world = 3
    procs = [ctx.Process(target=_rebalance_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(q, procs, 300)
    for rank, status, _ in res:
        assert status == "ok", f"rank {rank}: {status}"
    assert all(m >= 1 for _, _, m in res), "the boundaries never moved: the rebalance path was not exercised"
