// domain.hpp -- host-side state of one rank's device-resident chunk set (shared by domain.cu, peer.cu)
#pragma once

#include "common.cuh"

#include <memory>

namespace nixb200
{
struct Plan;   // peer.cu: which slabs cross to which rank
struct PeerCtx; // peer.cu: device tables, buffers and the NCCL communicator of the cross-rank exchange

struct Domain {
  nixb200_domain_desc     desc;
  Geo                     geo;
  std::vector<ChunkGeo>   cg_host;
  ChunkGeo*               cg_dev = nullptr;
  void*                   uf     = nullptr; // E/B of every chunk, in the domain's real type
  void*                   uj     = nullptr; // rho / J
  void*                   um     = nullptr; // moments [nchunk][Mz][My][Mx][ns][14], allocated on first use
  double*                 pack_dev = nullptr; // device staging of the packers' output (grown on demand)
  size_t                  pack_bytes = 0;
  int*                    count_dev = nullptr;
  std::vector<SpeciesDev> sp;
  cudaStream_t            stream = nullptr;
  // real type of the device-resident data: fp64 (the reference's, default) or fp32 (desc.fp32)
  bool                    fp32 = false;
  size_t                  esz  = 8; // bytes per real
  int                     nct  = 7; // reals per particle in the SoA store (fp32: the 64-bit id takes two)
  int                     fcs  = 6; // reals per E/B cell (fp32: 8, TMA strides are multiples of 16 bytes)
  std::vector<double>     origin_host;          // [nchunk][3] lower corner (z,y,x) of every chunk
  double*                 origin_dev = nullptr; // fp32 mode: particle positions are relative to it
  double*                 aos_tmp    = nullptr; // fp32 mode: fp64 AoS staging of the boundary (grown on demand)
  size_t                  aos_tmp_bytes = 0;
  bool                    owns_stream = false; // false once the caller has handed over its own stream
  bool                    has_remote  = false; // a neighbour chunk lives on another rank (needs set_ranks + comm)
  // capacity bookkeeping: after every sort the device reports, per species, (total particles, leaver
  // records, message particles) to pinned host memory; the next step grows the stores before they fill up
  int32_t*                stat_host    = nullptr; // pinned [ns][4]
  int32_t*                stat_dev     = nullptr; // device [ns][4]
  double*                 energy_dev   = nullptr; // device [nchunk][2]
  int64_t*                np_dev       = nullptr; // device [ns][nchunk]
  cudaEvent_t             ev_stat      = nullptr;
  bool                    stat_pending = false;
  CUtensorMap             tmap;
  void*                   scan_tmp = nullptr;
  int*                    err_dev  = nullptr;
  int*                    nbvalid_dev = nullptr;
  double*                 halo_buf    = nullptr; // device staging for the per-chunk buffer API
  size_t                  halo_buf_bytes = 0;
  cudaEvent_t             ev0 = nullptr, ev1 = nullptr;
  // overlapped host transfers (field_upload/download_overlapped): a second stream, per field (uf, uj)
  // the event after which the main stream has finished with the array for this step, and the event
  // after which the last overlapped copy of it is complete
  cudaStream_t            copy_stream = nullptr;
  cudaEvent_t             ev_main_done[2] = {nullptr, nullptr}, ev_copy_done[2] = {nullptr, nullptr};
  bool                    copy_pending[2] = {false, false};
  double*                 dense[2] = {nullptr, nullptr}; // interior-only staging of uf / uj (allocated on first use)
  bool                    timed = false;
  size_t                  cells_per_chunk = 0;
  bool                    particles_set   = false;
  bool                    profiling       = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending[NIXB200_NPHASE];
  std::vector<int>        coord_all; // id -> (cz,cy,cx) of every chunk of the box
  PeerCtx*                peer = nullptr; // non-null after nixb200_domain_set_ranks with neighbours on other ranks
  double                  phase_ms[NIXB200_NPHASE]    = {};
  int                     phase_calls[NIXB200_NPHASE] = {};
};

// RAII bracket: records an event pair around one phase when profiling is on
struct PhaseTimer {
  Domain*     d;
  int         phase;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  PhaseTimer(Domain* dd, int ph) : d(dd), phase(ph)
  {
    if (!d->profiling) return;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, d->stream);
  }
  ~PhaseTimer()
  {
    if (!e0) return;
    cudaEventRecord(e1, d->stream);
    d->pending[phase].push_back({e0, e1});
  }
};


// restores the caller's current device on scope exit (every entry point runs on the domain's device)
struct DeviceGuard {
  int  prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev)
  {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard()
  {
    if (switched) cudaSetDevice(prev);
  }
};

// domain.cu
int grow_particles(Domain* d, SpeciesDev& s, int64_t newcap);           // keeps xu
int grow_leavers(Domain* d, SpeciesDev& s, int64_t newlcap, bool keep); // keep: lrec survives
int record_stats(Domain* d);

int  alloc_species_particles(Domain* d, SpeciesDev& s, int64_t ntot, bool with_scratch = true); // (re)sizes the stores of one species
int  alloc_species_scratch(Domain* d, SpeciesDev& s);
void free_species_scratch(SpeciesDev& s);
int pack_scratch(Domain* d, size_t bytes);                           // d->pack_dev, grown on demand

// peer.cu: raw transport for rebalance.cu
void* peer_comm(const Domain* d);
bool  peer_take_comm(Domain* d, void** comm); // detach the communicator from d (returns whether d owned it)
int   peer_give_comm(Domain* d, void* comm, bool own);
int   peer_sendrecv_bytes(void* comm, cudaStream_t st, int n, const int* ranks, const void* const* sbuf, const size_t* sbytes,
                          void* const* rbuf, const size_t* rbytes);

// peer.cu
PeerTabs peer_tabs(const Domain* d);
void     peer_destroy(Domain* d);
int      peer_alloc_species(Domain* d, SpeciesDev& s);
int      peer_exchange_halo(Domain* d, int mode);       // pack -> NCCL send/recv (no-op without peers)
const void* peer_recvbuf(const Domain* d);
const void* peer_recvbuf_moment(const Domain* d);
int      peer_migrate(Domain* d);                       // the whole migrate + sort phase with peers
int      do_sort_species(Domain* d, SpeciesDev& s);
} // namespace nixb200
