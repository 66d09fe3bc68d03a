"""The CPU oracle split over several ranks (TEST INFRASTRUCTURE): every rank holds the oracle chunks
of its id range and moves halo buffers exactly as the reference does -- per chunk, per direction,
B.recv[e] <- neighbour(B, e).send[26 - e] (chunk.hpp:532-554) -- with `torch.distributed` (gloo)
standing in for MPI between ranks.  Which buffers cross, to whom and in which order comes from the
PRODUCT's plan (nix_b200.core.Plan, the host logic of csrc/peer.cu), so a run that matches the
single-process oracle domain proves the plan's routing."""
import numpy as np
import torch
import torch.distributed as dist

from nix_b200 import core
from oracle import nixoracle as no


class RankOracle:
    def __init__(self, lib, prob, boundary, rank):
        self.lib, self.prob, self.rank = lib, prob, rank
        self.boundary = np.asarray(boundary)
        self.ids = list(range(int(boundary[rank]), int(boundary[rank + 1])))
        cd = prob.cdims
        self.grid2id = {tuple(int(v) for v in c): i for i, c in enumerate(prob.coord)}
        gdims = tuple(cd[a] * prob.dims[a] for a in range(3))
        npr = [prob.ncell() * prob.ppc] * prob.ns
        self.chunks = {}
        for i in self.ids:
            off = tuple(int(prob.coord[i][a]) * prob.dims[a] for a in range(3))
            self.chunks[i] = no.Chunk(lib, prob.dims, prob.nb, prob.order, prob.ns, npr, prob.q, prob.m, off, gdims,
                                      prob.delh)
        self.plan = core.Plan(cd, prob.dims, prob.nb, prob.coord, boundary, rank)

    def neighbor(self, i, e):
        c, cd = self.prob.coord[i], self.prob.cdims
        d = (e // 9 - 1, (e // 3) % 3 - 1, e % 3 - 1)
        return self.grid2id[tuple((int(c[a]) + d[a]) % cd[a] for a in range(3))]

    def load(self):
        for i, c in self.chunks.items():
            c.uf[...] = self.prob.field(i)
            for s in range(self.prob.ns):
                c.set_particles(s, self.prob.particles(i, s))

    def sort_only(self):
        for c in self.chunks.values():
            for s in range(self.prob.ns):
                c.count(s, 0, c.np(s) - 1, True)
                c.sort(s)

    def exchange(self, mode):
        ch = self.chunks
        if mode == no.MODE_PARTICLE:
            for c in ch.values():
                for s in range(self.prob.ns):
                    c.count(s, 0, c.np(s) - 1, True)
        send = {}
        for i, c in ch.items():
            c.halo_pack(mode)
            size, addr = c.bufsize(mode).copy(), c.bufaddr(mode).copy()
            buf = c.sendbuf(mode).copy()
            send[i] = [bytes(buf[addr[d]:addr[d] + size[d]]) if d != 13 else b"" for d in range(27)]
        # messages to / from other ranks, in the plan's order (sizes first: they vary for particles)
        remote = {}
        for peer in self.plan.peers:
            r = peer["rank"]
            out = [send[i][d] for (i, d, _) in peer["send"]]
            osz = torch.tensor([len(b) for b in out], dtype=torch.int64)
            isz = torch.zeros(len(peer["recv"]), dtype=torch.int64)
            ops = [dist.P2POp(dist.isend, osz, r), dist.P2POp(dist.irecv, isz, r)]
            for q in dist.batch_isend_irecv(ops):
                q.wait()
            obuf = torch.frombuffer(bytearray(b"".join(out)) or bytearray(1), dtype=torch.uint8)
            ibuf = torch.zeros(max(1, int(isz.sum())), dtype=torch.uint8)
            ops = [dist.P2POp(dist.isend, obuf, r), dist.P2POp(dist.irecv, ibuf, r)]
            for q in dist.batch_isend_irecv(ops):
                q.wait()
            raw, pos = ibuf.numpy().tobytes(), 0
            for (i, e, _), n in zip(peer["recv"], isz.tolist()):
                remote[(i, e)] = raw[pos:pos + n]
                pos += n
        for i, c in ch.items():
            msgs = []
            for e in range(27):
                if e == 13:
                    msgs.append(b"")
                    continue
                nid = self.neighbor(i, e)
                msgs.append(send[nid][26 - e] if nid in ch else remote[(i, e)])
            c.set_recv_sizes(mode, [len(m) for m in msgs])
            addr = c.bufaddr(mode)
            rb = c.recvbuf(mode)
            for e, m in enumerate(msgs):
                if m:
                    rb[addr[e]:addr[e] + len(m)] = np.frombuffer(m, dtype=np.uint8)
            c.halo_unpack(mode)

    def rebalance(self, new_boundary):
        """Move to new rank boundaries the way nixb200_domain_rebalance does (host logic: nixb200_rebalance_moves):
        what leaves at the low end goes to rank-1, at the high end to rank+1 (Balancer::sendrecv_chunk,
        balancer.hpp:122-332); a chunk travels as its state between steps -- E/B, J and the cell-sorted particles
        of every species."""
        import pickle
        nrank = len(self.boundary) - 1
        new_boundary = np.asarray(new_boundary)
        old = (int(self.boundary[self.rank]), int(self.boundary[self.rank + 1]))
        new = (int(new_boundary[self.rank]), int(new_boundary[self.rank + 1]))
        rc, (sl, sr, rl, rr, keep) = core.rebalance_moves(old, new)
        assert rc == 0, "old and new range of a rank must overlap"
        prob = self.prob

        def state(i):
            c = self.chunks[i]
            return (i, c.uf.copy(), c.uj.copy(), [c.particles(s) for s in range(prob.ns)])

        for peer, send, recv in ((self.rank - 1, sl, rl), (self.rank + 1, sr, rr)):
            if peer < 0 or peer >= nrank:
                assert send[1] <= send[0] and recv[1] <= recv[0]
                continue
            out = pickle.dumps([state(i) for i in range(send[0], max(send))])
            osz, isz = torch.tensor([len(out)], dtype=torch.int64), torch.zeros(1, dtype=torch.int64)
            for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, osz, peer), dist.P2POp(dist.irecv, isz, peer)]):
                q.wait()
            obuf = torch.frombuffer(bytearray(out), dtype=torch.uint8)
            ibuf = torch.zeros(int(isz), dtype=torch.uint8)
            for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, obuf, peer), dist.P2POp(dist.irecv, ibuf, peer)]):
                q.wait()
            got = pickle.loads(ibuf.numpy().tobytes())
            assert [g[0] for g in got] == list(range(recv[0], max(recv)))
            cd = prob.cdims
            gdims = tuple(cd[a] * prob.dims[a] for a in range(3))
            for i, uf, uj, parts in got:
                off = tuple(int(prob.coord[i][a]) * prob.dims[a] for a in range(3))
                c = no.Chunk(self.lib, prob.dims, prob.nb, prob.order, prob.ns, [max(1, len(x)) for x in parts], prob.q,
                             prob.m, off, gdims, prob.delh)
                c.uf[...] = uf
                c.uj[...] = uj
                for s2, x in enumerate(parts):
                    c.set_particles(s2, x)
                self.chunks[i] = c
            for i in range(send[0], max(send)):
                del self.chunks[i]
        self.boundary = new_boundary
        self.ids = list(range(new[0], new[1]))
        assert sorted(self.chunks) == self.ids and (keep[1] - keep[0]) > 0
        self.plan = core.Plan(prob.cdims, prob.dims, prob.nb, prob.coord, new_boundary, self.rank)

    def step(self, delt, cc):
        for c in self.chunks.values():
            c.uj[...] = 0.0
            c.push_deposit(delt, cc)
        self.exchange(no.MODE_CURRENT)
        self.exchange(no.MODE_FIELD)
        self.exchange(no.MODE_PARTICLE)
