// peer.cu -- chunks partitioned over ranks (= GPUs) along the space-filling-curve order: which
// slabs cross a rank boundary (the plan), and their transport over NVLink with NCCL send/recv.
//
// Replaces, for neighbours that live on ANOTHER rank, the reference's per-chunk, per-direction
// MPI messages (26 x chunks x 3 modes per step):
//   Chunk::begin_bc_exchange / end_bc_exchange    chunk.hpp:507-586   (MPI_Isend / Irecv / Waitall)
//   Chunk::probe_bc_exchange                      chunk.cpp:310-395   (size discovery for particles)
//   ChunkMap::get_rank                            chunkmap.cpp:156-164 (owner of a chunk id)
// by ONE message per (peer rank, mode): every slab bound for a peer is packed into one contiguous
// device buffer in an order both sides derive independently -- sorted by (sender chunk id,
// direction) -- and moved with ncclSend/ncclRecv inside one group.  The receiving halo kernels
// (halo.cu) read the receive buffer exactly where they would have read a same-device neighbour, in
// the reference's unpack order, so results do not depend on the partition (bit-exact).
// Particle migration exchanges the per-slab counts first (the reference probes message sizes),
// then the payloads (56 bytes per particle, xtensor_halo3d.hpp:259).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): a single-rank domain never needs it, and
// the library keeps loading on machines without it.
#include "domain.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>

namespace nixb200
{
// -----------------------------------------------------------------------------------------------
// plan: pure host logic, no device needed (tested on CPU under gloo, tests/test_multirank_cpu.py)
// -----------------------------------------------------------------------------------------------
struct PlanEntry {
  int peer;       // rank on the other side
  int sender_id;  // global id of the SENDING chunk          } the canonical sort key,
  int sender_dir; // direction 0..26 the sender sends in     } identical on both sides
  int local_id;   // global id of MY chunk of this slab
  int local_dir;  // my direction (send list) / my receive slot (recv list) = 26 - sender_dir
  int cells;      // cells of the slab
};

struct Plan {
  int                    nrank = 1, rank = 0;
  int                    cdims[3], dims[3], nb = 0;
  std::vector<int>       boundary;
  std::vector<int>       peers; // ascending ranks I exchange with
  std::vector<PlanEntry> send, recv;
  std::vector<int>       send_first, recv_first; // [npeer+1]
};

static int rank_of(const std::vector<int>& boundary, int id)
{
  // ChunkMap::get_rank, chunkmap.cpp:156-164
  return (int)(std::upper_bound(boundary.begin(), boundary.end(), id) - boundary.begin()) - 1;
}

static int slab_cells(const int* dims, int nb, int dir)
{
  int e[3] = {dir / 9, (dir / 3) % 3, dir % 3};
  int n    = 1;
  for (int a = 0; a < 3; a++) n *= (e[a] == 1) ? dims[a] : nb;
  return n;
}

static Plan* plan_build(const int* cdims, const int* dims, int nb, const int* coord, int nrank,
                        const int* boundary, int rank)
{
  const int ncid = cdims[0] * cdims[1] * cdims[2];
  if (nrank < 1 || rank < 0 || rank >= nrank || boundary[0] != 0 || boundary[nrank] != ncid) {
    set_error("plan: boundary must run from 0 to the number of chunks, rank in [0, nrank)");
    return nullptr;
  }
  for (int r = 0; r < nrank; r++)
    if (boundary[r + 1] < boundary[r]) {
      set_error("plan: boundary must be ascending");
      return nullptr;
    }
  Plan* p  = new Plan();
  p->nrank = nrank;
  p->rank  = rank;
  p->nb    = nb;
  for (int a = 0; a < 3; a++) {
    p->cdims[a] = cdims[a];
    p->dims[a]  = dims[a];
  }
  p->boundary.assign(boundary, boundary + nrank + 1);
  std::vector<int> grid2id(ncid, -1);
  for (int id = 0; id < ncid; id++) {
    const int* c = &coord[3 * id];
    grid2id[(c[0] * cdims[1] + c[1]) * cdims[2] + c[2]] = id;
  }
  for (int id = boundary[rank]; id < boundary[rank + 1]; id++) {
    const int* c = &coord[3 * id];
    for (int d = 0; d < 27; d++) {
      if (d == 13) continue;
      int e[3] = {d / 9 - 1, (d / 3) % 3 - 1, d % 3 - 1};
      int n[3];
      for (int a = 0; a < 3; a++) n[a] = ((c[a] + e[a]) % cdims[a] + cdims[a]) % cdims[a];
      const int nid   = grid2id[(n[0] * cdims[1] + n[1]) * cdims[2] + n[2]];
      const int owner = rank_of(p->boundary, nid);
      if (owner == rank) continue;
      // I send my slab of direction d to `owner`; and the neighbour sends me ITS slab of direction
      // 26-d, which I receive in slot d (chunk.hpp:532-554)
      p->send.push_back({owner, id, d, id, d, slab_cells(dims, nb, d)});
      p->recv.push_back({owner, nid, 26 - d, id, d, slab_cells(dims, nb, d)});
    }
  }
  auto key = [](const PlanEntry& a, const PlanEntry& b) {
    if (a.peer != b.peer) return a.peer < b.peer;
    if (a.sender_id != b.sender_id) return a.sender_id < b.sender_id;
    return a.sender_dir < b.sender_dir;
  };
  std::sort(p->send.begin(), p->send.end(), key);
  std::sort(p->recv.begin(), p->recv.end(), key);
  for (auto& e : p->send)
    if (p->peers.empty() || p->peers.back() != e.peer) p->peers.push_back(e.peer);
  // (the recv list has the same peers: neighbourhood is symmetric)
  auto firsts = [&](const std::vector<PlanEntry>& v, std::vector<int>& first) {
    first.assign(p->peers.size() + 1, 0);
    size_t j = 0;
    for (size_t q = 0; q < p->peers.size(); q++) {
      first[q] = (int)j;
      while (j < v.size() && v[j].peer == p->peers[q]) j++;
    }
    first[p->peers.size()] = (int)v.size();
  };
  firsts(p->send, p->send_first);
  firsts(p->recv, p->recv_first);
  return p;
}

// -----------------------------------------------------------------------------------------------
// NCCL, resolved at run time
// -----------------------------------------------------------------------------------------------
struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                               = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                        = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t)                                                  = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)       = nullptr;
  ncclResult_t (*GroupStart)()                                                             = nullptr;
  ncclResult_t (*GroupEnd)()                                                               = nullptr;
  const char* (*GetErrorString)(ncclResult_t)                                              = nullptr;
};

static Nccl* nccl()
{
  static Nccl  n;
  static bool  tried = false;
  if (tried) return n.handle ? &n : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.handle) break;
  }
  if (!n.handle) {
    set_error(std::string("cannot load libnccl.so.2: ") + dlerror());
    return nullptr;
  }
  bool ok = true;
  auto sym = [&](const char* name) {
    void* f = dlsym(n.handle, name);
    if (!f) ok = false;
    return f;
  };
  n.GetUniqueId    = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
  n.CommInitRank   = (decltype(n.CommInitRank))sym("ncclCommInitRank");
  n.CommDestroy    = (decltype(n.CommDestroy))sym("ncclCommDestroy");
  n.Send           = (decltype(n.Send))sym("ncclSend");
  n.Recv           = (decltype(n.Recv))sym("ncclRecv");
  n.GroupStart     = (decltype(n.GroupStart))sym("ncclGroupStart");
  n.GroupEnd       = (decltype(n.GroupEnd))sym("ncclGroupEnd");
  n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
  if (!ok) {
    set_error("libnccl.so.2 lacks a required symbol");
    dlclose(n.handle);
    n.handle = nullptr;
    return nullptr;
  }
  return &n;
}

#define NIX_NCCL(call)                                                                           \
  do {                                                                                           \
    ncclResult_t r__ = (call);                                                                   \
    if (r__ != ncclSuccess) {                                                                    \
      set_error(std::string(#call) + ": " + nccl()->GetErrorString(r__));                        \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)

// -----------------------------------------------------------------------------------------------
// per-domain context of the cross-rank exchange
// -----------------------------------------------------------------------------------------------
struct PeerCtx {
  Plan*      plan = nullptr;
  ncclComm_t comm = nullptr;
  bool       own_comm = false;
  int        nsend = 0, nrecv = 0, ns = 0;
  // device tables
  PeerEntry* send_ent  = nullptr;
  PeerEntry* recv_ent  = nullptr;
  int32_t*   send_slot = nullptr;
  int32_t*   recv_slot = nullptr;
  // halo buffers, peer-major, sized for the 6-component field (the current uses 4/6 of them)
  double* hsend = nullptr;
  double* hrecv = nullptr;
  void*   msend = nullptr; // moment halo (ns * 14 values per cell), allocated on first use
  void*   mrecv = nullptr;
  std::vector<int64_t> send_cell_first, recv_cell_first; // [npeer+1] first cell of each peer
  // particle counts: per peer one message [ns][entries of the peer]
  int32_t* cnt_send = nullptr; // device [ns * nsend]
  int32_t* cnt_recv = nullptr; // device [ns * nrecv]
  int32_t* cnt_host = nullptr; // pinned [ns * (nsend + nrecv)]
  int32_t* tab_host = nullptr; // pinned [ns][(nsend+1) + (nrecv+1) + nrecv]
  std::vector<PeerEntry> send_host, recv_host;
  int64_t last_sent = 0, last_received = 0; // particles of the last migrate (all species)
};

PeerTabs peer_tabs(const Domain* d)
{
  PeerTabs t;
  std::memset(&t, 0, sizeof(t));
  if (d->peer) {
    t.send_ent  = d->peer->send_ent;
    t.recv_ent  = d->peer->recv_ent;
    t.send_slot = d->peer->send_slot;
    t.recv_slot = d->peer->recv_slot;
    t.nsend     = d->peer->nsend;
    t.nrecv     = d->peer->nrecv;
  }
  return t;
}

const void* peer_recvbuf(const Domain* d)
{
  return d->peer ? d->peer->hrecv : nullptr;
}

const void* peer_recvbuf_moment(const Domain* d)
{
  return d->peer ? d->peer->mrecv : nullptr;
}

void peer_destroy(Domain* d)
{
  PeerCtx* c = d->peer;
  if (!c) return;
  void* dev[] = {c->send_ent, c->recv_ent, c->send_slot, c->recv_slot, c->hsend, c->hrecv, c->cnt_send, c->cnt_recv, c->msend, c->mrecv};
  for (void* p : dev)
    if (p) cudaFree(p);
  if (c->cnt_host) cudaFreeHost(c->cnt_host);
  if (c->tab_host) cudaFreeHost(c->tab_host);
  if (c->comm && c->own_comm && nccl()) nccl()->CommDestroy(c->comm);
  delete c->plan;
  delete c;
  d->peer = nullptr;
}

static size_t ptab_ints(const PeerCtx* c)
{
  return (size_t)(c->nsend + 1) + (c->nrecv + 1) + c->nrecv;
}

int peer_alloc_species(Domain* d, SpeciesDev& s)
{
  PeerCtx* c = d->peer;
  if (!c) return 0;
  if (s.paysend) cudaFree(s.paysend);
  if (s.payrecv) cudaFree(s.payrecv);
  if (s.ptab) cudaFree(s.ptab);
  s.paysend = s.payrecv = nullptr;
  s.ptab               = nullptr;
  NIX_CUDA(cudaMalloc(&s.paysend, d->esz * d->nct * s.lcap));
  NIX_CUDA(cudaMalloc(&s.payrecv, d->esz * d->nct * s.lcap));
  NIX_CUDA(cudaMalloc(&s.ptab, sizeof(int32_t) * ptab_ints(c)));
  NIX_CUDA(cudaMemset(s.ptab, 0, sizeof(int32_t) * ptab_ints(c)));
  return 0;
}

// pack every slab bound for another rank, one ncclSend + ncclRecv per peer
int peer_exchange_halo(Domain* d, int mode)
{
  PeerCtx* c = d->peer;
  if (!c || c->plan->peers.empty()) return 0;
  if (!c->comm) {
    set_error("neighbours on other ranks but no communicator: call nixb200_domain_comm_init first");
    return 1;
  }
  // bytes per cell in the peer buffers = the device's own cell layout (E/B: 6 doubles or 8 floats; J: 4 reals;
  // moments: ns * 14 reals, in buffers of their own that exist only once moments were asked for)
  const int    ncm   = (int)d->sp.size() * 14;
  const size_t cellb = ((mode == NIXB200_MODE_FIELD) ? d->fcs : (mode == NIXB200_MODE_MOMENT ? ncm : 4)) * d->esz;
  const void*  data  = (mode == NIXB200_MODE_FIELD) ? d->uf : (mode == NIXB200_MODE_MOMENT ? d->um : d->uj);
  void *       sbuf = c->hsend, *rbuf = c->hrecv;
  if (mode == NIXB200_MODE_MOMENT) {
    if (!c->msend) {
      NIX_CUDA(cudaMalloc(&c->msend, cellb * std::max<int64_t>(1, c->send_cell_first.back())));
      NIX_CUDA(cudaMalloc(&c->mrecv, cellb * std::max<int64_t>(1, c->recv_cell_first.back())));
    }
    sbuf = c->msend, rbuf = c->mrecv;
  }
  if (launch_peer_pack(d->geo, mode, data, peer_tabs(d), sbuf, d->stream, d->fp32, ncm)) return 1;
  Nccl* n  = nccl();
  char* hs = reinterpret_cast<char*>(sbuf);
  char* hr = reinterpret_cast<char*>(rbuf);
  NIX_NCCL(n->GroupStart());
  for (size_t q = 0; q < c->plan->peers.size(); q++) {
    const int64_t s0 = c->send_cell_first[q], s1 = c->send_cell_first[q + 1];
    const int64_t r0 = c->recv_cell_first[q], r1 = c->recv_cell_first[q + 1];
    NIX_NCCL(n->Send(hs + s0 * cellb, (size_t)(s1 - s0) * cellb, ncclChar, c->plan->peers[q], c->comm, d->stream));
    NIX_NCCL(n->Recv(hr + r0 * cellb, (size_t)(r1 - r0) * cellb, ncclChar, c->plan->peers[q], c->comm, d->stream));
  }
  NIX_NCCL(n->GroupEnd());
  return 0;
}

// XtensorHaloParticle3D across ranks: counts -> (host learns the sizes) -> payloads -> append + sort
int peer_migrate(Domain* d)
{
  PeerCtx*   c  = d->peer;
  const Geo& g  = d->geo;
  const int  ns = (int)d->sp.size();
  PeerTabs   pt = peer_tabs(d);
  Nccl*      n  = nccl();
  const bool remote = !c->plan->peers.empty();
  if (remote && !c->comm) {
    set_error("neighbours on other ranks but no communicator: call nixb200_domain_comm_init first");
    return 1;
  }
  for (int is = 0; is < ns; is++) {
    if (launch_mig_scan(g, d->sp[is], d->stream)) return 1;
    if (launch_peer_counts(g, d->sp[is], pt, is, c->cnt_send, d->stream)) return 1;
  }
  const size_t tabn = ptab_ints(c);
  if (remote) {
    // 1. counts of every crossing slab, all species in one message per peer
    NIX_NCCL(n->GroupStart());
    for (size_t q = 0; q < c->plan->peers.size(); q++) {
      const int s0 = c->plan->send_first[q], s1 = c->plan->send_first[q + 1];
      const int r0 = c->plan->recv_first[q], r1 = c->plan->recv_first[q + 1];
      NIX_NCCL(n->Send(c->cnt_send + (size_t)ns * s0, (size_t)ns * (s1 - s0), ncclInt32, c->plan->peers[q], c->comm, d->stream));
      NIX_NCCL(n->Recv(c->cnt_recv + (size_t)ns * r0, (size_t)ns * (r1 - r0), ncclInt32, c->plan->peers[q], c->comm, d->stream));
    }
    NIX_NCCL(n->GroupEnd());
    // 2. the host needs the sizes to post the payload messages (the reference's probe, chunk.cpp:310-395)
    NIX_CUDA(cudaMemcpyAsync(c->cnt_host, c->cnt_send, sizeof(int32_t) * ns * c->nsend, cudaMemcpyDeviceToHost, d->stream));
    NIX_CUDA(cudaMemcpyAsync(c->cnt_host + (size_t)ns * c->nsend, c->cnt_recv, sizeof(int32_t) * ns * c->nrecv,
                             cudaMemcpyDeviceToHost, d->stream));
    NIX_CUDA(cudaStreamSynchronize(d->stream));
  }
  // 3. per species: payload offsets of every slab (peer-major, list order) and received counts
  c->last_sent = c->last_received = 0;
  std::vector<int64_t> nsent(ns, 0), nrecvd(ns, 0);
  for (int is = 0; is < ns; is++) {
    int32_t* tab   = c->tab_host + (size_t)is * tabn;
    int32_t* spoff = tab;
    int32_t* rpoff = tab + (c->nsend + 1);
    int32_t* rcnt  = rpoff + (c->nrecv + 1);
    int64_t  run   = 0;
    for (int j = 0; j < c->nsend; j++) {
      spoff[j] = (int32_t)run;
      run += c->cnt_host[c->send_host[j].cidx + is * c->send_host[j].cstride];
    }
    spoff[c->nsend] = (int32_t)run;
    nsent[is]       = run;
    run             = 0;
    for (int j = 0; j < c->nrecv; j++) {
      int v    = c->cnt_host[(size_t)ns * c->nsend + c->recv_host[j].cidx + is * c->recv_host[j].cstride];
      rpoff[j] = (int32_t)run;
      rcnt[j]  = v;
      run += v;
    }
    rpoff[c->nrecv] = (int32_t)run;
    nrecvd[is]      = run;
    // every rank must reach the payload group below with the sizes both sides agreed on: a buffer that
    // is too small is regrown here (the host knows the exact counts), never a reason to leave early
    if (nsent[is] > d->sp[is].lcap || nrecvd[is] > d->sp[is].lcap) {
      if (grow_leavers(d, d->sp[is], 2 * std::max(nsent[is], nrecvd[is]), true)) return 1;
    }
    c->last_sent += nsent[is];
    c->last_received += nrecvd[is];
    NIX_CUDA(cudaMemcpyAsync(d->sp[is].ptab, tab, sizeof(int32_t) * tabn, cudaMemcpyHostToDevice, d->stream));
    if (launch_mig_route(g, d->cg_dev, d->sp[is], pt, d->err_dev, d->stream, d->fp32)) return 1;
  }
  // 4. payloads: one message per (peer, species)
  if (remote) {
    NIX_NCCL(n->GroupStart());
    for (int is = 0; is < ns; is++) {
      const int32_t* tab   = c->tab_host + (size_t)is * tabn;
      const int32_t* spoff = tab;
      const int32_t* rpoff = tab + (c->nsend + 1);
      for (size_t q = 0; q < c->plan->peers.size(); q++) {
        const int64_t s0 = spoff[c->plan->send_first[q]], s1 = spoff[c->plan->send_first[q + 1]];
        const int64_t r0 = rpoff[c->plan->recv_first[q]], r1 = rpoff[c->plan->recv_first[q + 1]];
        const size_t pb = d->esz * d->nct; // bytes per particle (56 in fp64, xtensor_halo3d.hpp:259; 32 in fp32)
        if (s1 > s0)
          NIX_NCCL(n->Send(reinterpret_cast<char*>(d->sp[is].paysend) + s0 * pb, (size_t)(s1 - s0) * pb, ncclChar, c->plan->peers[q], c->comm, d->stream));
        if (r1 > r0)
          NIX_NCCL(n->Recv(reinterpret_cast<char*>(d->sp[is].payrecv) + r0 * pb, (size_t)(r1 - r0) * pb, ncclChar, c->plan->peers[q], c->comm, d->stream));
      }
    }
    NIX_NCCL(n->GroupEnd());
  }
  // 5. append what arrived, then count + sort
  for (int is = 0; is < ns; is++) {
    if (launch_mig_recv(g, d->cg_dev, d->sp[is], pt, (int)nrecvd[is], d->err_dev, d->stream, d->fp32)) return 1;
    if (do_sort_species(d, d->sp[is])) return 1;
  }
  return 0;
}

void* peer_comm(const Domain* d)
{
  return d->peer ? d->peer->comm : nullptr;
}

bool peer_take_comm(Domain* d, void** comm)
{
  *comm = nullptr;
  if (!d->peer) return false;
  *comm          = d->peer->comm;
  const bool own = d->peer->own_comm;
  d->peer->comm     = nullptr;
  d->peer->own_comm = false;
  return own;
}

int peer_give_comm(Domain* d, void* comm, bool own)
{
  if (!d->peer) {
    set_error("no rank partition on this domain");
    return 1;
  }
  d->peer->comm     = reinterpret_cast<ncclComm_t>(comm);
  d->peer->own_comm = own;
  return 0;
}

// n messages each way inside one NCCL group (a rank of MPI_PROC_NULL-like value < 0 is skipped)
int peer_sendrecv_bytes(void* comm, cudaStream_t st, int nmsg, const int* ranks, const void* const* sbuf, const size_t* sbytes,
                        void* const* rbuf, const size_t* rbytes)
{
  Nccl* n = nccl();
  if (!n || !comm) {
    set_error("no NCCL communicator");
    return 1;
  }
  NIX_NCCL(n->GroupStart());
  for (int q = 0; q < nmsg; q++) {
    if (ranks[q] < 0) continue;
    if (sbytes[q]) NIX_NCCL(n->Send(sbuf[q], sbytes[q], ncclChar, ranks[q], reinterpret_cast<ncclComm_t>(comm), st));
    if (rbytes[q]) NIX_NCCL(n->Recv(rbuf[q], rbytes[q], ncclChar, ranks[q], reinterpret_cast<ncclComm_t>(comm), st));
  }
  NIX_NCCL(n->GroupEnd());
  return 0;
}

static int peer_setup(Domain* d, Plan* plan)
{
  peer_destroy(d);
  PeerCtx* c = new PeerCtx();
  d->peer    = c;
  c->plan    = plan;
  c->nsend   = (int)plan->send.size();
  c->nrecv   = (int)plan->recv.size();
  c->ns      = (int)d->sp.size();
  const Geo& g     = d->geo;
  const int  begin = d->desc.id_begin;
  const int  ns    = c->ns;
  auto build = [&](const std::vector<PlanEntry>& v, const std::vector<int>& first, std::vector<PeerEntry>& out,
                   std::vector<int32_t>& slot, std::vector<int64_t>& cell_first) {
    out.resize(v.size());
    slot.assign((size_t)g.nchunk * 27, -1);
    cell_first.assign(plan->peers.size() + 1, 0);
    int64_t cell = 0;
    for (size_t q = 0; q < plan->peers.size(); q++) {
      cell_first[q] = cell;
      const int np  = first[q + 1] - first[q];
      for (int j = first[q]; j < first[q + 1]; j++) {
        out[j].k       = v[j].local_id - begin;
        out[j].dir     = v[j].local_dir;
        out[j].celloff = (int)cell;
        out[j].cells   = v[j].cells;
        out[j].cidx    = ns * first[q] + (j - first[q]);
        out[j].cstride = np;
        slot[(size_t)out[j].k * 27 + out[j].dir] = (int32_t)j;
        cell += v[j].cells;
      }
    }
    cell_first[plan->peers.size()] = cell;
    return cell;
  };
  std::vector<int32_t> sslot, rslot;
  int64_t scells = build(plan->send, plan->send_first, c->send_host, sslot, c->send_cell_first);
  int64_t rcells = build(plan->recv, plan->recv_first, c->recv_host, rslot, c->recv_cell_first);
  if (scells >= ((int64_t)1 << 31) || rcells >= ((int64_t)1 << 31)) {
    set_error("halo buffer exceeds 2^31 cells");
    return 1;
  }
  NIX_CUDA(cudaMalloc(&c->send_slot, sizeof(int32_t) * sslot.size()));
  NIX_CUDA(cudaMalloc(&c->recv_slot, sizeof(int32_t) * rslot.size()));
  NIX_CUDA(cudaMemcpy(c->send_slot, sslot.data(), sizeof(int32_t) * sslot.size(), cudaMemcpyHostToDevice));
  NIX_CUDA(cudaMemcpy(c->recv_slot, rslot.data(), sizeof(int32_t) * rslot.size(), cudaMemcpyHostToDevice));
  NIX_CUDA(cudaMalloc(&c->send_ent, sizeof(PeerEntry) * std::max(1, c->nsend)));
  NIX_CUDA(cudaMalloc(&c->recv_ent, sizeof(PeerEntry) * std::max(1, c->nrecv)));
  if (c->nsend) NIX_CUDA(cudaMemcpy(c->send_ent, c->send_host.data(), sizeof(PeerEntry) * c->nsend, cudaMemcpyHostToDevice));
  if (c->nrecv) NIX_CUDA(cudaMemcpy(c->recv_ent, c->recv_host.data(), sizeof(PeerEntry) * c->nrecv, cudaMemcpyHostToDevice));
  NIX_CUDA(cudaMalloc(&c->hsend, sizeof(double) * 6 * std::max<int64_t>(1, scells)));
  NIX_CUDA(cudaMalloc(&c->hrecv, sizeof(double) * 6 * std::max<int64_t>(1, rcells)));
  NIX_CUDA(cudaMalloc(&c->cnt_send, sizeof(int32_t) * std::max(1, ns * c->nsend)));
  NIX_CUDA(cudaMalloc(&c->cnt_recv, sizeof(int32_t) * std::max(1, ns * c->nrecv)));
  NIX_CUDA(cudaMallocHost(&c->cnt_host, sizeof(int32_t) * std::max(1, ns * (c->nsend + c->nrecv))));
  NIX_CUDA(cudaMallocHost(&c->tab_host, sizeof(int32_t) * ns * ptab_ints(c)));
  std::memset(c->cnt_host, 0, sizeof(int32_t) * std::max(1, ns * (c->nsend + c->nrecv)));
  for (auto& s : d->sp)
    if (peer_alloc_species(d, s)) return 1;
  return 0;
}
} // namespace nixb200

using namespace nixb200;

extern "C" {

int nixb200_plan_create(const int* cdims, const int* dims, int nb, const int* coord, int nrank,
                        const int* boundary, int rank, nixb200_plan** out)
{
  if (!cdims || !dims || !coord || !boundary || !out) {
    set_error("null argument");
    return 1;
  }
  Plan* p = plan_build(cdims, dims, nb, coord, nrank, boundary, rank);
  *out    = reinterpret_cast<nixb200_plan*>(p);
  return p ? 0 : 1;
}

int nixb200_plan_destroy(nixb200_plan* p)
{
  delete reinterpret_cast<Plan*>(p);
  return 0;
}

int nixb200_plan_npeer(const nixb200_plan* pp)
{
  const Plan* p = reinterpret_cast<const Plan*>(pp);
  return p ? (int)p->peers.size() : -1;
}

int nixb200_plan_peer(const nixb200_plan* pp, int q, int* peer_rank, int* nsend, int* nrecv)
{
  const Plan* p = reinterpret_cast<const Plan*>(pp);
  if (!p || q < 0 || q >= (int)p->peers.size() || !peer_rank || !nsend || !nrecv) {
    set_error("plan_peer: bad argument");
    return 1;
  }
  *peer_rank = p->peers[q];
  *nsend     = p->send_first[q + 1] - p->send_first[q];
  *nrecv     = p->recv_first[q + 1] - p->recv_first[q];
  return 0;
}

int nixb200_plan_entries(const nixb200_plan* pp, int q, int* send3, int* recv3)
{
  const Plan* p = reinterpret_cast<const Plan*>(pp);
  if (!p || q < 0 || q >= (int)p->peers.size()) {
    set_error("plan_entries: bad argument");
    return 1;
  }
  if (send3)
    for (int j = p->send_first[q], o = 0; j < p->send_first[q + 1]; j++, o++) {
      send3[3 * o + 0] = p->send[j].local_id;
      send3[3 * o + 1] = p->send[j].local_dir;
      send3[3 * o + 2] = p->send[j].cells;
    }
  if (recv3)
    for (int j = p->recv_first[q], o = 0; j < p->recv_first[q + 1]; j++, o++) {
      recv3[3 * o + 0] = p->recv[j].local_id;
      recv3[3 * o + 1] = p->recv[j].local_dir;
      recv3[3 * o + 2] = p->recv[j].cells;
    }
  return 0;
}

int nixb200_domain_set_ranks(nixb200_domain* dd, int nrank, const int* boundary, int rank)
{
  Domain* d = reinterpret_cast<Domain*>(dd);
  if (!d || !boundary) {
    set_error("null argument");
    return 1;
  }
  if (nrank < 1 || rank < 0 || rank >= nrank || boundary[rank] != d->desc.id_begin || boundary[rank + 1] != d->desc.id_end) {
    set_error("set_ranks: boundary[rank], boundary[rank+1] must equal the domain's id range");
    return 1;
  }
  DeviceGuard dev_guard(d->desc.device);
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  Plan* p = plan_build(d->desc.cdims, d->desc.dims, d->desc.nb, d->coord_all.data(), nrank, boundary, rank);
  if (!p) return 1;
  return peer_setup(d, p);
}

int nixb200_comm_unique_id(void* id128)
{
  Nccl* n = nccl();
  if (!n || !id128) return 1;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NIX_NCCL(n->GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof(id));
  return 0;
}

int nixb200_domain_comm_init(nixb200_domain* dd, const void* id128)
{
  Domain* d = reinterpret_cast<Domain*>(dd);
  if (!d || !id128 || !d->peer) {
    set_error("comm_init: call nixb200_domain_set_ranks first");
    return 1;
  }
  Nccl* n = nccl();
  if (!n) return 1;
  DeviceGuard dev_guard(d->desc.device);
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  NIX_NCCL(n->CommInitRank(&d->peer->comm, d->peer->plan->nrank, id, d->peer->plan->rank));
  d->peer->own_comm = true;
  return 0;
}

int nixb200_comm_create(int nrank, int rank, const void* id128, int device, void** comm_out)
{
  if (!id128 || !comm_out) {
    set_error("comm_create: null argument");
    return 1;
  }
  Nccl* n = nccl();
  if (!n) return 1;
  DeviceGuard  dev_guard(device);
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  NIX_NCCL(n->CommInitRank(&comm, nrank, id, rank));
  *comm_out = comm;
  return 0;
}

int nixb200_comm_destroy(void* comm)
{
  Nccl* n = nccl();
  if (!n || !comm) return 1;
  NIX_NCCL(n->CommDestroy(reinterpret_cast<ncclComm_t>(comm)));
  return 0;
}

int nixb200_device_count(int* n)
{
  if (!n) return 1;
  NIX_CUDA(cudaGetDeviceCount(n));
  return 0;
}

int nixb200_domain_set_comm(nixb200_domain* dd, void* nccl_comm)
{
  Domain* d = reinterpret_cast<Domain*>(dd);
  if (!d || !d->peer || !nccl_comm) {
    set_error("set_comm: call nixb200_domain_set_ranks first");
    return 1;
  }
  if (!nccl()) return 1;
  d->peer->comm     = reinterpret_cast<ncclComm_t>(nccl_comm);
  d->peer->own_comm = false;
  return 0;
}

int nixb200_domain_peer_traffic(nixb200_domain* dd, int64_t* halo_cells_sent, int64_t* particles_sent,
                                int64_t* particles_received)
{
  Domain* d = reinterpret_cast<Domain*>(dd);
  if (!d) return 1;
  int64_t cells = 0;
  if (d->peer && !d->peer->send_cell_first.empty()) cells = d->peer->send_cell_first.back();
  if (halo_cells_sent) *halo_cells_sent = cells;
  if (particles_sent) *particles_sent = d->peer ? d->peer->last_sent : 0;
  if (particles_received) *particles_received = d->peer ? d->peer->last_received : 0;
  return 0;
}

} // extern "C"
