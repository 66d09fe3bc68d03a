// domain.cu -- host side of libnixb200.so: the device-resident chunk set of one rank and the
// extern "C" layer declared in include/nixb200.h.
//
// A domain owns, for a contiguous range of chunk ids (ChunkMap::get_rank, chunkmap.cpp:156-164):
//   uf [nchunk][Mz][My][Mx][6], uj [nchunk][Mz][My][Mx][4]      -- each chunk with its own ghosts,
//                                                                  exactly the reference's arrays
//   per species: SoA particles of ALL chunks concatenated in chunk order (xu / xv double buffer),
//                the global (chunk, cell, lane) histogram and its scan (= pcount / pindex).
#include "domain.hpp"

#include <algorithm>
#include <cstring>
#include <mutex>

namespace nixb200
{
static thread_local std::string g_error;
std::atomic<int64_t>            g_launches{0};

void set_error(const std::string& msg)
{
  g_error = msg;
}

static int required_nb(int order)
{
  return (order == 3) ? 3 : 2; // stencil extents of gather and deposit, DESIGN.md section 2
}

static int make_tensor_map(Domain* d)
{
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);
  void*                            fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  NIX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return 1;
  }
  const Geo&  g       = d->geo;
  // one cell = 6 doubles = 48 bytes, or 8 floats = 32 bytes (global strides must be multiples of 16 bytes)
  const cuuint64_t fc = (cuuint64_t)d->fcs, cb = fc * d->esz;
  cuuint64_t  dims[5] = {fc, (cuuint64_t)g.M[2], (cuuint64_t)g.M[1], (cuuint64_t)g.M[0], (cuuint64_t)g.nchunk};
  cuuint64_t  strides[4] = {cb, (cuuint64_t)g.M[2] * cb, (cuuint64_t)g.M[1] * g.M[2] * cb,
                            (cuuint64_t)g.M[0] * g.M[1] * g.M[2] * cb};
  int ez, ey, ex;
  push_tile_box(g.order, ez, ey, ex); // compile-time box of the push kernel (stencil box + bank padding)
  cuuint32_t  box[5]  = {(cuuint32_t)fc, (cuuint32_t)ex, (cuuint32_t)ey, (cuuint32_t)ez, 1};
  cuuint32_t  estr[5] = {1, 1, 1, 1, 1};
  CUresult    r = ((encode_fn)fn)(&d->tmap, d->fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, d->uf, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 1;
  }
  return 0;
}

static int free_species(SpeciesDev& s)
{
  void* ptrs[] = {s.xu,      s.xv,     s.key,    s.ordl,  s.hist,   s.start, s.oob,    s.cbase, s.cbase_new,
                  s.slabcnt, s.sendcnt, s.msgoff, s.recvoff, s.nleave, s.nmsg, s.lrec,   s.msg,   s.msgkey, s.paysend, s.payrecv, s.ptab};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  std::memset(&s, 0, sizeof(s));
  return 0;
}

static int alloc_species_fixed(Domain* d, SpeciesDev& s)
{
  const Geo&   g = d->geo;
  const size_t n = (size_t)g.nchunk * g.ncell * LANES;
  NIX_CUDA(cudaMalloc(&s.hist, sizeof(int32_t) * n));
  NIX_CUDA(cudaMalloc(&s.start, sizeof(int32_t) * (n + 4)));
  NIX_CUDA(cudaMalloc(&s.oob, sizeof(int32_t) * g.nchunk * LANES));
  NIX_CUDA(cudaMalloc(&s.cbase, sizeof(int32_t) * (g.nchunk + 1)));
  NIX_CUDA(cudaMalloc(&s.cbase_new, sizeof(int32_t) * (g.nchunk + 1)));
  NIX_CUDA(cudaMalloc(&s.slabcnt, sizeof(int32_t) * (size_t)g.nchunk * g.slaboff[27]));
  NIX_CUDA(cudaMalloc(&s.sendcnt, sizeof(int32_t) * g.nchunk * 27));
  NIX_CUDA(cudaMalloc(&s.msgoff, sizeof(int32_t) * g.nchunk * 27));
  NIX_CUDA(cudaMalloc(&s.recvoff, sizeof(int32_t) * g.nchunk * 27));
  NIX_CUDA(cudaMalloc(&s.nleave, sizeof(int32_t)));
  NIX_CUDA(cudaMalloc(&s.nmsg, sizeof(int32_t)));
  NIX_CUDA(cudaMemset(s.hist, 0, sizeof(int32_t) * n));
  NIX_CUDA(cudaMemset(s.start, 0, sizeof(int32_t) * (n + 4)));
  NIX_CUDA(cudaMemset(s.oob, 0, sizeof(int32_t) * g.nchunk * LANES));
  NIX_CUDA(cudaMemset(s.cbase, 0, sizeof(int32_t) * (g.nchunk + 1)));
  NIX_CUDA(cudaMemset(s.cbase_new, 0, sizeof(int32_t) * (g.nchunk + 1)));
  NIX_CUDA(cudaMemset(s.slabcnt, 0, sizeof(int32_t) * (size_t)g.nchunk * g.slaboff[27]));
  NIX_CUDA(cudaMemset(s.sendcnt, 0, sizeof(int32_t) * g.nchunk * 27));
  NIX_CUDA(cudaMemset(s.nleave, 0, sizeof(int32_t)));
  NIX_CUDA(cudaMemset(s.nmsg, 0, sizeof(int32_t)));
  return 0;
}

static int64_t round_cap(int64_t n)
{
  return (n + 127) / 128 * 128;
}

static int alloc_leavers(Domain* d, SpeciesDev& s, int64_t lcap)
{
  s.lcap = lcap;
  NIX_CUDA(cudaMalloc(&s.lrec, sizeof(int4) * lcap));
  NIX_CUDA(cudaMalloc(&s.msg, d->esz * d->nct * lcap));
  NIX_CUDA(cudaMalloc(&s.msgkey, sizeof(int32_t) * lcap));
  NIX_CUDA(cudaMemset(s.msgkey, 0xff, sizeof(int32_t) * lcap));
  if (d->peer && peer_alloc_species(d, s)) return 1;
  return 0;
}

// everything of a species that carries no state between steps (the temporary array, keys, member lists,
// migration buffers): a rebalance frees it first and re-creates it last to keep its peak memory down
void free_species_scratch(SpeciesDev& s)
{
  void* old[] = {s.xv, s.key, s.ordl, s.lrec, s.msg, s.msgkey, s.paysend, s.payrecv};
  for (void* p : old)
    if (p) cudaFree(p);
  s.xv = nullptr;
  s.key = s.ordl = nullptr;
  s.lrec = nullptr;
  s.msg = nullptr;
  s.msgkey = nullptr;
  s.paysend = s.payrecv = nullptr;
}

int alloc_species_scratch(Domain* d, SpeciesDev& s)
{
  free_species_scratch(s);
  NIX_CUDA(cudaMalloc(&s.xv, d->esz * d->nct * s.cap));
  NIX_CUDA(cudaMalloc(&s.key, sizeof(int32_t) * s.cap));
  NIX_CUDA(cudaMalloc(&s.ordl, sizeof(int32_t) * s.cap));
  NIX_CUDA(cudaMemset(s.xv, 0, d->esz * d->nct * s.cap));
  return alloc_leavers(d, s, round_cap(std::max<int64_t>(4096, s.cap / 16))); // regrown on demand (ensure_capacity)
}

int alloc_species_particles(Domain* d, SpeciesDev& s, int64_t ntot, bool with_scratch)
{
  free_species_scratch(s);
  if (s.xu) cudaFree(s.xu);
  s.xu = nullptr;
  double f = d->desc.capacity_factor > 0 ? d->desc.capacity_factor : 1.25;
  if (f < 1.0) f = 1.0;
  int64_t cap = round_cap((int64_t)((double)ntot * f) + 1024);
  if (cap >= (int64_t)1 << 31) {
    set_error("more than 2^31 particles of one species on one device");
    return 1;
  }
  s.cap = cap;
  NIX_CUDA(cudaMalloc(&s.xu, d->esz * d->nct * cap));
  NIX_CUDA(cudaMemset(s.xu, 0, d->esz * d->nct * cap));
  return with_scratch ? alloc_species_scratch(d, s) : 0;
}

// fp32 mode: fp64 AoS scratch for the particle boundary
static int aos_scratch(Domain* d, size_t bytes)
{
  if (bytes <= d->aos_tmp_bytes) return 0;
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  if (d->aos_tmp) cudaFree(d->aos_tmp);
  d->aos_tmp = nullptr, d->aos_tmp_bytes = 0;
  NIX_CUDA(cudaMalloc(&d->aos_tmp, bytes));
  d->aos_tmp_bytes = bytes;
  return 0;
}

static int check_chunk(Domain* d, int k)
{
  if (!d) {
    set_error("null domain");
    return 1;
  }
  if (k < 0 || k >= d->geo.nchunk) {
    set_error("chunk index out of range");
    return 1;
  }
  return 0;
}

static int check_species(Domain* d, int is)
{
  if (!d) {
    set_error("null domain");
    return 1;
  }
  if (is < 0 || is >= (int)d->sp.size()) {
    set_error("species index out of range");
    return 1;
  }
  return 0;
}

int do_sort_species(Domain* d, SpeciesDev& s)
{
  if (launch_sort(d->geo, d->cg_dev, s, d->err_dev, d->scan_tmp, d->stream, d->fp32)) return 1;
  std::swap(s.xu, s.xv);           // XtensorParticle::swap, xtensor_particle.hpp:120-123
  std::swap(s.cbase, s.cbase_new); // Np = pindex(Ng), xtensor_particle.hpp:320
  return 0;
}

// The reference's containers grow on demand (XtensorParticle::resize from pre_unpack,
// xtensor_halo3d.hpp:464-476).  Here the stores are sized once and regrown from the host BEFORE they fill
// up: only xu carries state between steps (xv, key, ordl are scratch), and the move runs on the domain's
// stream.
int grow_particles(Domain* d, SpeciesDev& s, int64_t newcap)
{
  newcap = round_cap(newcap);
  if (newcap <= s.cap) return 0;
  if (newcap >= (int64_t)1 << 31) {
    set_error("more than 2^31 particles of one species on one device");
    return 1;
  }
  double*  nxu = nullptr;
  double*  nxv = nullptr;
  int32_t *nkey = nullptr, *nordl = nullptr;
  NIX_CUDA(cudaMalloc(&nxu, d->esz * d->nct * newcap));
  NIX_CUDA(cudaMalloc(&nxv, d->esz * d->nct * newcap));
  NIX_CUDA(cudaMalloc(&nkey, sizeof(int32_t) * newcap));
  NIX_CUDA(cudaMalloc(&nordl, sizeof(int32_t) * newcap));
  NIX_CUDA(cudaMemsetAsync(nxu, 0, d->esz * d->nct * newcap, d->stream));
  NIX_CUDA(cudaMemsetAsync(nxv, 0, d->esz * d->nct * newcap, d->stream));
  for (int c = 0; c < d->nct; c++)
    NIX_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(nxu) + (size_t)c * newcap * d->esz,
                             reinterpret_cast<char*>(s.xu) + (size_t)c * s.cap * d->esz, d->esz * s.cap,
                             cudaMemcpyDeviceToDevice, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  cudaFree(s.xu);
  cudaFree(s.xv);
  cudaFree(s.key);
  cudaFree(s.ordl);
  s.xu = nxu, s.xv = nxv, s.key = nkey, s.ordl = nordl;
  s.cap = newcap;
  return 0;
}

int grow_leavers(Domain* d, SpeciesDev& s, int64_t newlcap, bool keep)
{
  newlcap = round_cap(newlcap);
  if (newlcap <= s.lcap) return 0;
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  int4*         olrec = s.lrec;
  const int64_t olcap = s.lcap;
  cudaFree(s.msg);
  cudaFree(s.msgkey);
  s.lrec = nullptr, s.msg = nullptr, s.msgkey = nullptr;
  if (alloc_leavers(d, s, newlcap)) return 1; // (the peer payload buffers follow s.lcap)
  if (keep) NIX_CUDA(cudaMemcpy(s.lrec, olrec, sizeof(int4) * olcap, cudaMemcpyDeviceToDevice));
  cudaFree(olrec);
  return 0;
}

// after a sort: (particles, leaver records, message particles) of every species -> pinned host memory
int record_stats(Domain* d)
{
  const int ns = (int)d->sp.size();
  for (int is = 0; is < ns; is++)
    if (launch_stats(d->geo, d->sp[is], d->stat_dev + 4 * is, d->stream)) return 1;
  NIX_CUDA(cudaMemcpyAsync(d->stat_host, d->stat_dev, sizeof(int32_t) * 4 * ns, cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaEventRecord(d->ev_stat, d->stream));
  d->stat_pending = true;
  return 0;
}

// before a push: look at what the last sort reported and make room.  The wait is for work that was
// enqueued a whole step ago.
static int ensure_capacity(Domain* d)
{
  if (!d->stat_pending) return 0;
  NIX_CUDA(cudaEventSynchronize(d->ev_stat));
  d->stat_pending = false;
  for (size_t is = 0; is < d->sp.size(); is++) {
    SpeciesDev&    s  = d->sp[is];
    const int32_t* st = d->stat_host + 4 * is;
    const int64_t  total = st[0], moving = std::max(st[1], st[2]);
    if (total > s.cap || moving > s.lcap) {
      set_error("particle store overflowed during the last step (raise capacity_factor or call nixb200_domain_reserve): "
                "the particles of this domain are lost");
      return 1;
    }
    // headroom for one more step: what moved last step may arrive on top
    if (total + 2 * moving + 1024 > s.cap && grow_particles(d, s, (int64_t)(1.25 * (double)(total + 2 * moving)) + 4096)) return 1;
    if (2 * moving > s.lcap && grow_leavers(d, s, 4 * moving, false)) return 1;
  }
  return 0;
}
// device staging shared by the packers and the chunk wire format
int pack_scratch(Domain* d, size_t bytes)
{
  if (!d->count_dev) NIX_CUDA(cudaMalloc(&d->count_dev, sizeof(int)));
  if (bytes <= d->pack_bytes) return 0;
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  if (d->pack_dev) cudaFree(d->pack_dev);
  d->pack_dev = nullptr, d->pack_bytes = 0;
  NIX_CUDA(cudaMalloc(&d->pack_dev, bytes));
  d->pack_bytes = bytes;
  return 0;
}

} // namespace nixb200

using namespace nixb200;

static Domain* D(nixb200_domain* d)
{
  return reinterpret_cast<Domain*>(d);
}

// every entry point that takes a domain runs with the domain's device current and restores the
// caller's device on return (several domains on several devices may live in one process / thread)
#define NIX_ENTER(dd)                                                                            \
  Domain* d = D(dd);                                                                             \
  if (!d) {                                                                                      \
    set_error("null domain");                                                                    \
    return 1;                                                                                    \
  }                                                                                              \
  DeviceGuard dev_guard__(d->desc.device)

static int need_peers(Domain* d)
{
  if (d->has_remote && !d->peer) {
    set_error("a neighbour chunk lives on another rank: call nixb200_domain_set_ranks (and comm_init) first");
    return 1;
  }
  return 0;
}

extern "C" {

const char* nixb200_last_error(void)
{
  return g_error.c_str();
}

const char* nixb200_version(void)
{
  return "nixb200 0.1 (sm_100a)";
}

int64_t nixb200_launch_count(void)
{
  return g_launches;
}

int nixb200_domain_create(const nixb200_domain_desc* desc, const int* coord, const double* q,
                          const double* m, nixb200_domain** out)
{
  if (!desc || !coord || !q || !m || !out) {
    set_error("null argument");
    return 1;
  }
  *out = nullptr;
  // the descriptor first (host only), then the device: a bad descriptor is reported as such on any machine
  if (desc->pusher < NIXB200_PUSH_BORIS || desc->pusher > NIXB200_PUSH_HIGUERA_CARY) {
    set_error("pusher must be NIXB200_PUSH_BORIS, _VAY or _HIGUERA_CARY");
    return 1;
  }
  if (desc->order < 1 || desc->order > 3) {
    set_error("order must be 1, 2 or 3");
    return 1;
  }
  if (desc->nb < required_nb(desc->order)) {
    set_error("boundary margin too small for this order (need 2 for order 1-2, 3 for order 3)");
    return 1;
  }
  if (desc->ns < 1 || desc->id_end <= desc->id_begin) {
    set_error("empty domain");
    return 1;
  }
  for (int a = 0; a < 3; a++) {
    if (desc->dims[a] < desc->nb || desc->cdims[a] < 1) {
      set_error("chunk dims must be >= boundary margin and cdims >= 1");
      return 1;
    }
    if (!(desc->del[a] > 0.0)) {
      set_error("cell sizes must be positive");
      return 1;
    }
  }
  if (!(desc->cc > 0.0)) {
    set_error("the speed of light must be positive");
    return 1;
  }
  if (desc->id_begin < 0 || desc->id_end > desc->cdims[0] * desc->cdims[1] * desc->cdims[2]) {
    set_error("chunk id range outside the box");
    return 1;
  }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: libnixb200 has no CPU fallback");
    return 1;
  }
  DeviceGuard dev_guard(desc->device);
  int         cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != desc->device) {
    set_error("cannot select CUDA device " + std::to_string(desc->device));
    return 1;
  }
  cudaDeviceProp prop;
  NIX_CUDA(cudaGetDeviceProperties(&prop, desc->device));
  if (prop.major != 10) {
    set_error(std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
              "; this library is built for sm_100a only");
    return 1;
  }
  Domain* d = new Domain();
  d->desc   = *desc;
  d->fp32   = desc->fp32 != 0;
  d->esz    = d->fp32 ? 4 : 8;
  d->nct    = d->fp32 ? 8 : 7;
  d->fcs    = d->fp32 ? 8 : 6;
  Geo& g    = d->geo;
  std::memset(&g, 0, sizeof(g));
  g.nb     = desc->nb;
  g.order  = desc->order;
  g.is_odd = desc->order % 2;
  g.half   = desc->order / 2;
  g.nchunk = desc->id_end - desc->id_begin;
  for (int a = 0; a < 3; a++) {
    g.N[a]    = desc->dims[a];
    g.M[a]    = desc->dims[a] + 2 * desc->nb;
    g.R[a]    = desc->dims[a] + 1;
    g.nc[a]   = g.R[a];
    g.del[a]  = desc->del[a];
    g.rdel[a] = 1 / desc->del[a];
    // chunk.cpp:220-236 and xtensor_particle.hpp:361-369
    g.glo[a]  = 0.0;
    g.ghi[a]  = (desc->cdims[a] * desc->dims[a]) * desc->del[a];
    g.glen[a] = 1 * (g.ghi[a] - g.glo[a]);
  }
  g.cc    = desc->cc;
  g.rc    = 1 / desc->cc;
  g.ncell = g.R[0] * g.R[1] * g.R[2];
  if (choose_push_tile(g)) {
    delete d;
    set_error("no push tile fits shared memory");
    return 1;
  }
  g.slabt = 2;
  {
    int off = 0;
    for (int s = 0; s < 27; s++) {
      g.slaboff[s] = off;
      if (s == 13) continue;
      int e[3] = {s / 9, (s / 3) % 3, s % 3};
      int n    = 1;
      for (int a = 0; a < 3; a++) n *= (e[a] == 1) ? g.nc[a] : g.slabt;
      off += n;
    }
    g.slaboff[27] = off;
  }
  if ((double)g.nchunk * g.ncell * LANES >= 2147483000.0) {
    delete d;
    set_error("key space (nchunk*ncell*8) exceeds int32");
    return 1;
  }

  // neighbour table from the id -> coordinate map (ChunkVector::set_neighbors, chunkvector.hpp:56-81;
  // periodic neighbour coordinates, chunkmap.cpp:142-154)
  const int* cd  = desc->cdims;
  const int  ncid = cd[0] * cd[1] * cd[2];
  std::vector<int> grid2id(ncid, -1);
  for (int id = 0; id < ncid; id++) {
    const int* c = &coord[3 * id];
    if (c[0] < 0 || c[0] >= cd[0] || c[1] < 0 || c[1] >= cd[1] || c[2] < 0 || c[2] >= cd[2]) {
      delete d;
      set_error("chunk coordinate out of range");
      return 1;
    }
    grid2id[(c[0] * cd[1] + c[1]) * cd[2] + c[2]] = id;
  }
  d->coord_all.assign(coord, coord + 3 * ncid);
  d->cg_host.resize(g.nchunk);
  for (int k = 0; k < g.nchunk; k++) {
    const int* c  = &coord[3 * (desc->id_begin + k)];
    ChunkGeo&  cg = d->cg_host[k];
    for (int a = 0; a < 3; a++) {
      int    off = c[a] * g.N[a];
      double del = g.del[a];
      d->origin_host.push_back(off * del);
      if (d->fp32) off = 0;                    // fp32 mode: positions are relative to the chunk's origin
      cg.lo[a]   = off * del;                  // chunk.cpp:217,224,231
      cg.hi[a]   = off * del + g.N[a] * del;   // chunk.cpp:218,225,232
      cg.off[a]  = cg.lo[a] - 0.5 * del * g.is_odd;       // xtensor_particle.hpp:332-334
      cg.hoff[a] = cg.lo[a] - 0.5 * del * (1 - g.is_odd); // ref_driver.cpp
      cg.imin[a] = cg.lo[a] + 0.5 * del;
    }
    for (int s = 0; s < 27; s++) {
      int e[3] = {s / 9 - 1, (s / 3) % 3 - 1, s % 3 - 1};
      int n[3];
      for (int a = 0; a < 3; a++) n[a] = ((c[a] + e[a]) % cd[a] + cd[a]) % cd[a];
      int id = grid2id[(n[0] * cd[1] + n[1]) * cd[2] + n[2]];
      if (id < 0) cg.nbr[s] = -1;
      else if (id >= desc->id_begin && id < desc->id_end) cg.nbr[s] = id - desc->id_begin;
      else {
        cg.nbr[s]     = -2; // lives on another rank
        d->has_remote = true;
      }
    }
  }

  d->cells_per_chunk = (size_t)g.M[0] * g.M[1] * g.M[2];
  auto fail = [&](const char* what) {
    set_error(std::string(what) + ": " + cudaGetErrorString(cudaGetLastError()));
    nixb200_domain_destroy(reinterpret_cast<nixb200_domain*>(d));
    return 1;
  };
  if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
  d->owns_stream = true;
  if (push_deposit_prepare(g.order, d->fp32)) {
    std::string e = g_error;
    nixb200_domain_destroy(reinterpret_cast<nixb200_domain*>(d));
    set_error(e);
    return 1;
  }
  if (cudaMalloc(&d->cg_dev, sizeof(ChunkGeo) * g.nchunk) != cudaSuccess) return fail("cudaMalloc");
  if (cudaMemcpy(d->cg_dev, d->cg_host.data(), sizeof(ChunkGeo) * g.nchunk, cudaMemcpyHostToDevice) != cudaSuccess)
    return fail("cudaMemcpy");
  if (cudaMalloc(&d->uf, d->esz * d->fcs * d->cells_per_chunk * g.nchunk) != cudaSuccess) return fail("cudaMalloc uf");
  if (cudaMalloc(&d->uj, d->esz * 4 * d->cells_per_chunk * g.nchunk) != cudaSuccess) return fail("cudaMalloc uj");
  cudaMemset(d->uf, 0, d->esz * d->fcs * d->cells_per_chunk * g.nchunk);
  cudaMemset(d->uj, 0, d->esz * 4 * d->cells_per_chunk * g.nchunk);
  if (cudaMalloc(&d->origin_dev, sizeof(double) * 3 * g.nchunk) != cudaSuccess) return fail("cudaMalloc origin");
  cudaMemcpy(d->origin_dev, d->origin_host.data(), sizeof(double) * 3 * g.nchunk, cudaMemcpyHostToDevice);
  if (cudaMalloc(&d->scan_tmp, scan_tmp_bytes((size_t)g.nchunk * g.ncell * LANES)) != cudaSuccess) return fail("cudaMalloc scan");
  if (cudaMalloc(&d->err_dev, 2 * sizeof(int)) != cudaSuccess) return fail("cudaMalloc err");
  cudaMemset(d->err_dev, 0, 2 * sizeof(int));
  if (cudaMalloc(&d->stat_dev, sizeof(int32_t) * 4 * desc->ns) != cudaSuccess) return fail("cudaMalloc stat");
  if (cudaMallocHost(&d->stat_host, sizeof(int32_t) * 4 * desc->ns) != cudaSuccess) return fail("cudaMallocHost stat");
  cudaEventCreateWithFlags(&d->ev_stat, cudaEventDisableTiming);
  if (cudaMalloc(&d->nbvalid_dev, sizeof(int) * 27) != cudaSuccess) return fail("cudaMalloc");
  d->halo_buf_bytes = sizeof(double) * 6 * d->cells_per_chunk;
  if (cudaMalloc(&d->halo_buf, d->halo_buf_bytes) != cudaSuccess) return fail("cudaMalloc halo");
  cudaEventCreate(&d->ev0);
  cudaEventCreate(&d->ev1);
  if (cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return fail("copy stream");
  for (int w = 0; w < 2; w++) {
    cudaEventCreateWithFlags(&d->ev_main_done[w], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&d->ev_copy_done[w], cudaEventDisableTiming);
  }

  d->sp.resize(desc->ns);
  for (int is = 0; is < desc->ns; is++) {
    std::memset(&d->sp[is], 0, sizeof(SpeciesDev));
    d->sp[is].q = q[is];
    d->sp[is].m = m[is];
    if (alloc_species_fixed(d, d->sp[is]) || alloc_species_particles(d, d->sp[is], 0)) {
      std::string e = g_error;
      nixb200_domain_destroy(reinterpret_cast<nixb200_domain*>(d));
      set_error(e);
      return 1;
    }
  }
  if (make_tensor_map(d)) {
    std::string e = g_error;
    nixb200_domain_destroy(reinterpret_cast<nixb200_domain*>(d));
    set_error(e);
    return 1;
  }
  *out = reinterpret_cast<nixb200_domain*>(d);
  return 0;
}

int nixb200_domain_destroy(nixb200_domain* dd)
{
  Domain* d = D(dd);
  if (!d) return 0;
  DeviceGuard dev_guard(d->desc.device);
  cudaDeviceSynchronize();
  peer_destroy(d);
  for (auto& s : d->sp) free_species(s);
  if (d->cg_dev) cudaFree(d->cg_dev);
  if (d->uf) cudaFree(d->uf);
  if (d->uj) cudaFree(d->uj);
  if (d->scan_tmp) cudaFree(d->scan_tmp);
  if (d->err_dev) cudaFree(d->err_dev);
  if (d->stat_dev) cudaFree(d->stat_dev);
  if (d->energy_dev) cudaFree(d->energy_dev);
  if (d->np_dev) cudaFree(d->np_dev);
  if (d->origin_dev) cudaFree(d->origin_dev);
  if (d->um) cudaFree(d->um);
  if (d->pack_dev) cudaFree(d->pack_dev);
  if (d->count_dev) cudaFree(d->count_dev);
  if (d->aos_tmp) cudaFree(d->aos_tmp);
  if (d->stat_host) cudaFreeHost(d->stat_host);
  if (d->ev_stat) cudaEventDestroy(d->ev_stat);
  if (d->nbvalid_dev) cudaFree(d->nbvalid_dev);
  if (d->halo_buf) cudaFree(d->halo_buf);
  if (d->ev0) cudaEventDestroy(d->ev0);
  if (d->ev1) cudaEventDestroy(d->ev1);
  for (int w = 0; w < 2; w++) {
    if (d->ev_main_done[w]) cudaEventDestroy(d->ev_main_done[w]);
    if (d->ev_copy_done[w]) cudaEventDestroy(d->ev_copy_done[w]);
  }
  for (int w = 0; w < 2; w++)
    if (d->dense[w]) cudaFree(d->dense[w]);
  if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
  if (d->stream && d->owns_stream) cudaStreamDestroy(d->stream); // never a stream the caller handed over
  delete d;
  return 0;
}

int nixb200_domain_set_stream(nixb200_domain* dd, void* cuda_stream)
{
  NIX_ENTER(dd);
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  if (d->owns_stream) cudaStreamDestroy(d->stream); // the library's own stream; a caller's stream is never destroyed
  d->stream      = reinterpret_cast<cudaStream_t>(cuda_stream);
  d->owns_stream = false;
  return 0;
}

int nixb200_domain_synchronize(nixb200_domain* dd)
{
  NIX_ENTER(dd);
  if (!d) return 1;
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  return 0;
}

int nixb200_domain_check(nixb200_domain* dd, int* errbits)
{
  NIX_ENTER(dd);
  if (!errbits) return 1;
  int e[2] = {0, 0};
  NIX_CUDA(cudaMemcpyAsync(e, d->err_dev, 2 * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaMemsetAsync(d->err_dev, 0, sizeof(int), d->stream)); // the overflow flag e[1] stays until new particles are set
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  *errbits = e[0] | (e[1] ? NIXB200_ERR_CAPACITY : 0);
  return 0;
}

// address of chunk k inside uf / uj (bytes: the device arrays hold the domain's real type)
static char* chunk_array(Domain* d, int which, int k)
{
  const int fc = (which == NIXB200_FIELD_UF) ? d->fcs : 4;
  char*     b  = reinterpret_cast<char*>((which == NIXB200_FIELD_UF) ? d->uf : d->uj);
  return b + (size_t)k * d->cells_per_chunk * fc * d->esz;
}

// fp32 mode: one chunk's fp64 array <-> the device floats, through the fp64 staging buffer of the domain
static int chunk_convert(Domain* d, int which, int k, bool to_dev)
{
  const int nc = (which == NIXB200_FIELD_UF) ? 6 : 4, fc = (which == NIXB200_FIELD_UF) ? d->fcs : 4;
  return launch_cells_convert(to_dev, d->halo_buf, reinterpret_cast<float*>(chunk_array(d, which, k)), d->cells_per_chunk, nc,
                              fc, d->stream);
}

int nixb200_chunk_field_upload(nixb200_domain* dd, int k, int which, const double* host)
{
  NIX_ENTER(dd);
  if (check_chunk(d, k) || !host) return 1;
  const int    nc    = (which == NIXB200_FIELD_UF) ? 6 : 4;
  const size_t bytes = sizeof(double) * nc * d->cells_per_chunk;
  if (d->fp32) {
    NIX_CUDA(cudaMemcpyAsync(d->halo_buf, host, bytes, cudaMemcpyHostToDevice, d->stream));
    if (chunk_convert(d, which, k, true)) return 1;
  } else {
    NIX_CUDA(cudaMemcpyAsync(chunk_array(d, which, k), host, bytes, cudaMemcpyHostToDevice, d->stream));
  }
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  return 0;
}

int nixb200_chunk_field_download(nixb200_domain* dd, int k, int which, double* host)
{
  NIX_ENTER(dd);
  if (check_chunk(d, k) || !host) return 1;
  const int    nc    = (which == NIXB200_FIELD_UF) ? 6 : 4;
  const size_t bytes = sizeof(double) * nc * d->cells_per_chunk;
  if (d->fp32) {
    if (chunk_convert(d, which, k, false)) return 1;
    NIX_CUDA(cudaMemcpyAsync(host, d->halo_buf, bytes, cudaMemcpyDeviceToHost, d->stream));
  } else {
    NIX_CUDA(cudaMemcpyAsync(host, chunk_array(d, which, k), bytes, cudaMemcpyDeviceToHost, d->stream));
  }
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  return 0;
}

int nixb200_domain_field_upload_async(nixb200_domain* dd, int which, const double* host)
{
  NIX_ENTER(dd);
  if (!host) return 1;
  const int    nc    = (which == NIXB200_FIELD_UF) ? 6 : 4;
  const size_t bytes = sizeof(double) * nc * d->cells_per_chunk;
  if (d->fp32) { // chunk by chunk through the staging buffer (stream order keeps it consistent)
    for (int k = 0; k < d->geo.nchunk; k++) {
      NIX_CUDA(cudaMemcpyAsync(d->halo_buf, reinterpret_cast<const char*>(host) + (size_t)k * bytes, bytes, cudaMemcpyHostToDevice, d->stream));
      if (chunk_convert(d, which, k, true)) return 1;
    }
    return 0;
  }
  NIX_CUDA(cudaMemcpyAsync(chunk_array(d, which, 0), host, bytes * d->geo.nchunk, cudaMemcpyHostToDevice, d->stream));
  return 0;
}

int nixb200_domain_field_download_async(nixb200_domain* dd, int which, double* host)
{
  NIX_ENTER(dd);
  if (!host) return 1;
  const int    nc    = (which == NIXB200_FIELD_UF) ? 6 : 4;
  const size_t bytes = sizeof(double) * nc * d->cells_per_chunk;
  if (d->fp32) {
    for (int k = 0; k < d->geo.nchunk; k++) {
      if (chunk_convert(d, which, k, false)) return 1;
      NIX_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(host) + (size_t)k * bytes, d->halo_buf, bytes, cudaMemcpyDeviceToHost, d->stream));
    }
    return 0;
  }
  NIX_CUDA(cudaMemcpyAsync(host, chunk_array(d, which, 0), bytes * d->geo.nchunk, cudaMemcpyDeviceToHost, d->stream));
  return 0;
}

// ---- overlapped transfers: the copy runs on the domain's second stream, ordered against the phases
//      that touch the array (uf: push_deposit, exchange_field; uj: clear_current, push_deposit,
//      exchange_current) by events, so that e.g. the download of J and the upload of the next E/B
//      run while the main stream is still migrating and sorting particles ------------------------
static int field_index(int which) { return which == NIXB200_FIELD_UF ? 0 : 1; }

// main-stream phases call this before touching the array
static int wait_copy(Domain* d, int w)
{
  if (d->copy_pending[w]) {
    NIX_CUDA(cudaStreamWaitEvent(d->stream, d->ev_copy_done[w], 0));
    d->copy_pending[w] = false;
  }
  return 0;
}
// ... and this after their last access of the step
static int mark_main_done(Domain* d, int w)
{
  NIX_CUDA(cudaEventRecord(d->ev_main_done[w], d->stream));
  return 0;
}

static int field_copy_overlapped(Domain* d, int which, double* host, bool upload, bool interior)
{
  if (!d || !host) return 1;
  DeviceGuard  dev_guard(d->desc.device);
  const int    w     = field_index(which);
  const int    nc    = (w == 0) ? 6 : 4;
  const Geo&   g     = d->geo;
  const size_t cells = interior ? (size_t)g.N[0] * g.N[1] * g.N[2] : d->cells_per_chunk;
  const size_t bytes = sizeof(double) * nc * cells * g.nchunk;
  double*      full  = reinterpret_cast<double*>((w == 0) ? d->uf : d->uj);
  double*      dev   = full;
  if (d->fp32 && !interior) {
    set_error("fp32 mode: the full-array overlapped transfers are fp64 only (use the interior variants or the per-chunk calls)");
    return 1;
  }
  if (interior) {
    if (!d->dense[w]) NIX_CUDA(cudaMalloc(&d->dense[w], bytes));
    dev = d->dense[w];
  }
  auto interior_kernel = [&](bool pack) {
    if (d->fp32) return launch_interior_f32(pack, reinterpret_cast<float*>(full), dev, g, nc, (w == 0) ? d->fcs : 4, d->copy_stream);
    return launch_interior(pack, full, dev, g, nc, d->copy_stream);
  };
  NIX_CUDA(cudaStreamWaitEvent(d->copy_stream, d->ev_main_done[w], 0)); // no-op before the first phase
  if (upload) {
    NIX_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, d->copy_stream));
    if (interior && interior_kernel(false)) return 1;
  } else {
    if (interior && interior_kernel(true)) return 1;
    NIX_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, d->copy_stream));
  }
  NIX_CUDA(cudaEventRecord(d->ev_copy_done[w], d->copy_stream));
  d->copy_pending[w] = true;
  return 0;
}

int nixb200_domain_field_upload_overlapped(nixb200_domain* dd, int which, const double* host)
{
  return field_copy_overlapped(D(dd), which, const_cast<double*>(host), true, false);
}

int nixb200_domain_interior_upload_overlapped(nixb200_domain* dd, int which, const double* host)
{
  return field_copy_overlapped(D(dd), which, const_cast<double*>(host), true, true);
}

int nixb200_domain_interior_download_overlapped(nixb200_domain* dd, int which, double* host)
{
  return field_copy_overlapped(D(dd), which, host, false, true);
}

int nixb200_domain_field_download_overlapped(nixb200_domain* dd, int which, double* host)
{
  return field_copy_overlapped(D(dd), which, host, false, false);
}

int nixb200_domain_copy_synchronize(nixb200_domain* dd)
{
  NIX_ENTER(dd);
  if (!d) return 1;
  NIX_CUDA(cudaStreamSynchronize(d->copy_stream));
  return 0;
}

int nixb200_domain_set_profiling(nixb200_domain* dd, int on)
{
  NIX_ENTER(dd);
  if (!d) return 1;
  d->profiling = on != 0;
  return 0;
}

int nixb200_domain_get_phase_ms(nixb200_domain* dd, int phase, double* ms_sum, int* calls)
{
  NIX_ENTER(dd);
  if (!d || phase < 0 || phase >= NIXB200_NPHASE || !ms_sum || !calls) return 1;
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  for (auto& pr : d->pending[phase]) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) {
      d->phase_ms[phase] += t;
      d->phase_calls[phase]++;
    }
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  d->pending[phase].clear();
  *ms_sum               = d->phase_ms[phase];
  *calls                = d->phase_calls[phase];
  d->phase_ms[phase]    = 0;
  d->phase_calls[phase] = 0;
  return 0;
}

int nixb200_domain_set_particles(nixb200_domain* dd, int is, const double* xu_aos, const int64_t* np_chunk)
{
  NIX_ENTER(dd);
  if (check_species(d, is) || !np_chunk) return 1;
  const Geo&           g = d->geo;
  std::vector<int32_t> cbase(g.nchunk + 1, 0);
  int64_t              ntot = 0;
  for (int k = 0; k < g.nchunk; k++) {
    if (np_chunk[k] < 0) {
      set_error("negative particle count");
      return 1;
    }
    ntot += np_chunk[k];
    if (ntot >= ((int64_t)1 << 31) - 4096) {
      set_error("more than 2^31 particles of one species on one device");
      return 1;
    }
    cbase[k + 1] = (int32_t)ntot;
  }
  SpeciesDev& s = d->sp[is];
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  d->stat_pending = false;
  NIX_CUDA(cudaMemsetAsync(d->err_dev + 1, 0, sizeof(int), d->stream)); // fresh particles: the overflow flag goes
  if (alloc_species_particles(d, s, ntot)) return 1;
  NIX_CUDA(cudaMemcpyAsync(s.cbase, cbase.data(), sizeof(int32_t) * (g.nchunk + 1), cudaMemcpyHostToDevice, d->stream));
  if (ntot > 0) {
    if (!xu_aos) {
      set_error("null particle array");
      return 1;
    }
    if (d->fp32) {
      // fp64 AoS staged in a scratch buffer, converted into the fp32 store (positions relative to the chunk)
      if (aos_scratch(d, sizeof(double) * NC * ntot)) return 1;
      NIX_CUDA(cudaMemcpyAsync(d->aos_tmp, xu_aos, sizeof(double) * NC * ntot, cudaMemcpyDefault, d->stream));
      const double extent[3] = {g.N[0] * g.del[0], g.N[1] * g.del[1], g.N[2] * g.del[2]};
      if (launch_aos_to_soa_f32(d->aos_tmp, reinterpret_cast<float*>(s.xu), s.cap, 0, (size_t)ntot, s.cbase, g.nchunk,
                                d->origin_dev, extent, d->stream))
        return 1;
    } else {
      // AoS staged in the xv buffer (same byte size), transposed into xu; the source may be host or device memory
      NIX_CUDA(cudaMemcpyAsync(s.xv, xu_aos, sizeof(double) * NC * ntot, cudaMemcpyDefault, d->stream));
      if (launch_aos_to_soa(s.xv, s.xu, s.cap, 0, (size_t)ntot, d->stream)) return 1;
    }
  }
  // start[] of an unsorted container: only the chunk bases are meaningful until domain_sort()
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  d->particles_set = true;
  return 0;
}

int nixb200_domain_get_np(nixb200_domain* dd, int is, int64_t* np_chunk)
{
  NIX_ENTER(dd);
  if (check_species(d, is) || !np_chunk) return 1;
  const Geo&           g = d->geo;
  std::vector<int32_t> cbase(g.nchunk + 1);
  NIX_CUDA(cudaMemcpyAsync(cbase.data(), d->sp[is].cbase, sizeof(int32_t) * (g.nchunk + 1), cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  for (int k = 0; k < g.nchunk; k++) np_chunk[k] = cbase[k + 1] - cbase[k];
  return 0;
}

int nixb200_chunk_get_particles(nixb200_domain* dd, int k, int is, double* xu_aos, int64_t max_np, int64_t* np)
{
  NIX_ENTER(dd);
  if (check_chunk(d, k) || check_species(d, is) || !np) return 1;
  SpeciesDev& s = d->sp[is];
  int32_t     cb[2];
  NIX_CUDA(cudaMemcpyAsync(cb, s.cbase + k, sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  int64_t n = cb[1] - cb[0];
  *np       = n;
  if (!xu_aos || n == 0) return 0;
  if (n > max_np) {
    set_error("output buffer too small");
    return 1;
  }
  if (d->fp32) {
    if (aos_scratch(d, sizeof(double) * NC * n)) return 1;
    if (launch_soa_to_aos_f32(reinterpret_cast<const float*>(s.xu), d->aos_tmp, s.cap, (size_t)cb[0], (size_t)n,
                              d->origin_dev + 3 * k, d->stream))
      return 1;
    NIX_CUDA(cudaMemcpyAsync(xu_aos, d->aos_tmp, sizeof(double) * NC * n, cudaMemcpyDeviceToHost, d->stream));
    NIX_CUDA(cudaStreamSynchronize(d->stream));
    return 0;
  }
  // xv is scratch between steps (the reference's "temporary particle array", xtensor_particle.hpp:16)
  if (launch_soa_to_aos(s.xu, s.xv, s.cap, (size_t)cb[0], (size_t)n, d->stream)) return 1;
  NIX_CUDA(cudaMemcpyAsync(xu_aos, s.xv, sizeof(double) * NC * n, cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  return 0;
}

// expand the device's compact rows [0, ncell) (+ out-of-bounds row) to the reference's [Ng+1] layout
int nixb200_chunk_get_pindex(nixb200_domain* dd, int k, int is, int32_t* pindex)
{
  NIX_ENTER(dd);
  if (check_chunk(d, k) || check_species(d, is) || !pindex) return 1;
  const Geo&           g = d->geo;
  SpeciesDev&          s = d->sp[is];
  const size_t         n = (size_t)g.ncell * LANES;
  std::vector<int32_t> st(n + 1);
  NIX_CUDA(cudaMemcpyAsync(st.data(), s.start + (size_t)k * n, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  const int Ng   = (int)d->cells_per_chunk;
  const int base = st[0];
  const int npk  = st[n] - base;
  for (int ii = 0; ii <= Ng; ii++) pindex[ii] = (ii < g.ncell) ? st[(size_t)ii * LANES] - base : npk;
  return 0;
}

int nixb200_chunk_get_pcount(nixb200_domain* dd, int k, int is, int32_t* pcount)
{
  NIX_ENTER(dd);
  if (check_chunk(d, k) || check_species(d, is) || !pcount) return 1;
  const Geo&           g = d->geo;
  SpeciesDev&          s = d->sp[is];
  const size_t         n = (size_t)g.ncell * LANES;
  std::vector<int32_t> st(n + 1);
  int32_t              oob[LANES];
  NIX_CUDA(cudaMemcpyAsync(st.data(), s.start + (size_t)k * n, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaMemcpyAsync(oob, s.oob + (size_t)k * LANES, sizeof(int32_t) * LANES, cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  const int Ng = (int)d->cells_per_chunk;
  std::memset(pcount, 0, sizeof(int32_t) * LANES * ((size_t)Ng + 1));
  for (size_t j = 0; j < n; j++) pcount[j] = st[j + 1] - st[j];
  for (int l = 0; l < LANES; l++) pcount[(size_t)Ng * LANES + l] = oob[l];
  return 0;
}

int nixb200_domain_sort(nixb200_domain* dd)
{
  NIX_ENTER(dd);
  if (!d) return 1;
  PhaseTimer pt(d, 4);
  for (auto& s : d->sp) {
    if (launch_count_only(d->geo, d->cg_dev, s, d->err_dev, d->stream, d->fp32)) return 1;
    if (do_sort_species(d, s)) return 1;
  }
  return record_stats(d);
}

int nixb200_domain_clear_current(nixb200_domain* dd)
{
  NIX_ENTER(dd);
  if (!d) return 1;
  if (wait_copy(d, 1)) return 1;
  NIX_CUDA(cudaMemsetAsync(d->uj, 0, d->esz * 4 * d->cells_per_chunk * d->geo.nchunk, d->stream));
  return 0;
}

int nixb200_domain_push_deposit(nixb200_domain* dd, double delt)
{
  NIX_ENTER(dd);
  if (!d) return 1;
  if (ensure_capacity(d)) return 1;
  if (wait_copy(d, 0) || wait_copy(d, 1)) return 1;
  PhaseTimer pt(d, 0);
  NIX_CUDA(cudaEventRecord(d->ev0, d->stream));
  for (auto& s : d->sp) {
    PushArgs a;
    a.geo  = d->geo;
    a.cg   = d->cg_dev;
    a.uf   = d->uf;
    a.uj   = d->uj;
    a.sp   = s;
    a.delt = delt;
    a.err  = d->err_dev;
    a.pusher = d->desc.pusher;
    a.fp32   = d->fp32;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (d->profiling) {
      for (auto& e : ev) cudaEventCreate(&e);
      d->pending[5].push_back({ev[0], ev[1]});
      d->pending[6].push_back({ev[2], ev[3]});
    }
    if (launch_push_deposit(a, &d->tmap, d->desc.strict_fp != 0, d->stream, d->profiling ? ev : nullptr)) return 1;
  }
  NIX_CUDA(cudaEventRecord(d->ev1, d->stream));
  d->timed = true;
  return mark_main_done(d, 0); // the push is a reader of uf: an overlapped upload must wait for it
}

int nixb200_domain_exchange_current(nixb200_domain* dd)
{
  NIX_ENTER(dd);
  if (!d) return 1;
  if (need_peers(d) || wait_copy(d, 1)) return 1;
  {
    PhaseTimer pt(d, 1);
    if (peer_exchange_halo(d, NIXB200_MODE_CURRENT)) return 1;
    if (launch_halo_current(d->geo, d->cg_dev, d->uj, peer_tabs(d), peer_recvbuf(d), d->stream, d->fp32)) return 1;
  }
  return mark_main_done(d, 1);
}

int nixb200_domain_exchange_field(nixb200_domain* dd)
{
  NIX_ENTER(dd);
  if (!d) return 1;
  if (need_peers(d) || wait_copy(d, 0)) return 1;
  {
    PhaseTimer pt(d, 2);
    if (peer_exchange_halo(d, NIXB200_MODE_FIELD)) return 1;
    if (launch_halo_field(d->geo, d->cg_dev, d->uf, peer_tabs(d), peer_recvbuf(d), d->stream, d->fp32)) return 1;
  }
  return mark_main_done(d, 0);
}

int nixb200_domain_migrate_sort(nixb200_domain* dd)
{
  NIX_ENTER(dd);
  if (!d) return 1;
  if (need_peers(d)) return 1;
  PhaseTimer pt(d, 3);
  if (d->peer) return peer_migrate(d) || record_stats(d);
  const PeerTabs none = peer_tabs(d);
  for (auto& s : d->sp) {
    if (launch_mig_scan(d->geo, s, d->stream)) return 1;
    if (launch_mig_route(d->geo, d->cg_dev, s, none, d->err_dev, d->stream, d->fp32)) return 1;
    if (do_sort_species(d, s)) return 1;
  }
  return record_stats(d);
}

int nixb200_domain_step(nixb200_domain* dd, double delt)
{
  if (nixb200_domain_clear_current(dd)) return 1;
  if (nixb200_domain_push_deposit(dd, delt)) return 1;
  if (nixb200_domain_exchange_current(dd)) return 1;
  if (nixb200_domain_exchange_field(dd)) return 1;
  if (nixb200_domain_migrate_sort(dd)) return 1;
  return 0;
}

int nixb200_domain_push_bfd(nixb200_domain* dd, double delt, int ext)
{
  NIX_ENTER(dd);
  if (wait_copy(d, 0)) return 1;
  {
    PhaseTimer pt(d, 7);
    if (launch_push_bfd(d->geo, d->uf, delt, ext, d->stream, d->fp32)) return 1;
  }
  return mark_main_done(d, 0);
}

int nixb200_domain_push_efd(nixb200_domain* dd, double delt, double cfj)
{
  NIX_ENTER(dd);
  if (wait_copy(d, 0) || wait_copy(d, 1)) return 1;
  {
    PhaseTimer pt(d, 7);
    if (launch_push_efd(d->geo, d->uf, d->uj, delt, cfj, d->stream, d->fp32)) return 1;
  }
  return mark_main_done(d, 0) || mark_main_done(d, 1);
}

int nixb200_domain_step_em(nixb200_domain* dd, double delt, double cfj)
{
  if (nixb200_domain_clear_current(dd)) return 1;
  if (nixb200_domain_push_deposit(dd, delt)) return 1;
  if (nixb200_domain_exchange_current(dd)) return 1;
  if (nixb200_domain_push_bfd(dd, 0.5 * delt, 1)) return 1;
  if (nixb200_domain_push_efd(dd, delt, cfj)) return 1;
  if (nixb200_domain_exchange_field(dd)) return 1;
  if (nixb200_domain_push_bfd(dd, 0.5 * delt, 0)) return 1;
  if (nixb200_domain_exchange_field(dd)) return 1;
  if (nixb200_domain_migrate_sort(dd)) return 1;
  return 0;
}

int nixb200_domain_field_energy(nixb200_domain* dd, double* host_e2b2)
{
  NIX_ENTER(dd);
  if (!host_e2b2) return 1;
  if (wait_copy(d, 0)) return 1;
  if (!d->energy_dev) NIX_CUDA(cudaMalloc(&d->energy_dev, sizeof(double) * 2 * d->geo.nchunk));
  if (launch_field_energy(d->geo, d->uf, d->energy_dev, d->stream, d->fp32)) return 1;
  NIX_CUDA(cudaMemcpyAsync(host_e2b2, d->energy_dev, sizeof(double) * 2 * d->geo.nchunk, cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  return 0;
}

// History diagnostics without a host round trip per step: field energies [nchunk][2] and particle counts
// [ns][nchunk] (int64) are written to caller-provided PINNED host memory in stream order; the caller reads them
// after nixb200_domain_synchronize (or any later synchronising call).
int nixb200_domain_history_async(nixb200_domain* dd, double* host_e2b2, int64_t* host_np)
{
  NIX_ENTER(dd);
  if (host_e2b2) {
    if (wait_copy(d, 0)) return 1;
    if (!d->energy_dev) NIX_CUDA(cudaMalloc(&d->energy_dev, sizeof(double) * 2 * d->geo.nchunk));
    if (launch_field_energy(d->geo, d->uf, d->energy_dev, d->stream, d->fp32)) return 1;
    NIX_CUDA(cudaMemcpyAsync(host_e2b2, d->energy_dev, sizeof(double) * 2 * d->geo.nchunk, cudaMemcpyDeviceToHost, d->stream));
  }
  if (host_np) {
    const int nch = d->geo.nchunk;
    if (!d->np_dev) NIX_CUDA(cudaMalloc(&d->np_dev, sizeof(int64_t) * nch * d->sp.size()));
    for (size_t is = 0; is < d->sp.size(); is++)
      if (launch_np_from_cbase(d->sp[is].cbase, nch, d->np_dev + is * nch, d->stream)) return 1;
    NIX_CUDA(cudaMemcpyAsync(host_np, d->np_dev, sizeof(int64_t) * nch * d->sp.size(), cudaMemcpyDeviceToHost, d->stream));
  }
  return 0;
}

int nixb200_domain_set_strict_fp(nixb200_domain* dd, int on)
{
  NIX_ENTER(dd);
  d->desc.strict_fp = on ? 1 : 0;
  return 0;
}

// ---- diagnostics / output (diag.cu) ---------------------------------------------------------------------
static size_t moment_chunk_values(const Domain* d)
{
  return d->cells_per_chunk * d->sp.size() * 14;
}

int nixb200_domain_deposit_moment(nixb200_domain* dd)
{
  NIX_ENTER(dd);
  if (need_peers(d)) return 1;
  const size_t bytes = moment_chunk_values(d) * d->geo.nchunk * d->esz;
  if (!d->um) NIX_CUDA(cudaMalloc(&d->um, bytes));
  NIX_CUDA(cudaMemsetAsync(d->um, 0, bytes, d->stream));
  const int ns = (int)d->sp.size();
  for (int is = 0; is < ns; is++)
    if (launch_moment(d->geo, d->cg_dev, d->sp[is], is, ns, d->um, d->stream, d->fp32)) return 1;
  if (peer_exchange_halo(d, NIXB200_MODE_MOMENT)) return 1;
  return launch_halo_moment(d->geo, d->cg_dev, d->um, ns * 14, peer_tabs(d), peer_recvbuf_moment(d), d->stream, d->fp32);
}

int nixb200_chunk_moment_download(nixb200_domain* dd, int k, double* host)
{
  NIX_ENTER(dd);
  if (check_chunk(d, k) || !host) return 1;
  if (!d->um) {
    set_error("no moments on the device: call nixb200_domain_deposit_moment first");
    return 1;
  }
  const size_t n = moment_chunk_values(d);
  if (d->fp32) {
    if (pack_scratch(d, sizeof(double) * n)) return 1;
    if (launch_cells_convert(false, d->pack_dev, reinterpret_cast<float*>(d->um) + (size_t)k * n, n, 1, 1, d->stream)) return 1;
    NIX_CUDA(cudaMemcpyAsync(host, d->pack_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, d->stream));
  } else {
    NIX_CUDA(cudaMemcpyAsync(host, reinterpret_cast<double*>(d->um) + (size_t)k * n, sizeof(double) * n, cudaMemcpyDeviceToHost, d->stream));
  }
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  return 0;
}

static int pack_grid(Domain* d, int k, const void* base, bool colocate, int decimate, int nc, int fc, double* host, int64_t* count)
{
  if (decimate < 1) {
    set_error("decimate must be >= 1");
    return 1;
  }
  const int n = pack_count(d->geo, decimate, nc);
  if (count) *count = n;
  if (!host) return 0;
  if (pack_scratch(d, sizeof(double) * n)) return 1;
  const char* chunk = reinterpret_cast<const char*>(base) + (size_t)k * d->cells_per_chunk * fc * d->esz;
  if (launch_pack_grid(d->geo, chunk, colocate, decimate, nc, fc, d->pack_dev, d->stream, d->fp32)) return 1;
  NIX_CUDA(cudaMemcpyAsync(host, d->pack_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  return 0;
}

int nixb200_chunk_pack_field(nixb200_domain* dd, int k, int decimate, double* host, int64_t* count)
{
  NIX_ENTER(dd);
  if (check_chunk(d, k)) return 1;
  return pack_grid(d, k, d->uf, true, decimate, 6, d->fcs, host, count);
}

int nixb200_chunk_pack_moment(nixb200_domain* dd, int k, int which, int decimate, double* host, int64_t* count)
{
  NIX_ENTER(dd);
  if (check_chunk(d, k)) return 1;
  if (which == 0) return pack_grid(d, k, d->uj, false, decimate, 4, 4, host, count);
  if (!d->um) {
    set_error("no moments on the device: call nixb200_domain_deposit_moment first");
    return 1;
  }
  const int nc = (int)d->sp.size() * 14;
  return pack_grid(d, k, d->um, false, decimate, nc, nc, host, count);
}

int nixb200_chunk_pack_tracer(nixb200_domain* dd, int k, int is, double* host_aos, int64_t max_np, int64_t* np)
{
  NIX_ENTER(dd);
  if (check_chunk(d, k) || check_species(d, is) || !np) return 1;
  SpeciesDev& s = d->sp[is];
  int32_t     cb[2];
  NIX_CUDA(cudaMemcpyAsync(cb, s.cbase + k, sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  const int n = cb[1] - cb[0];
  const int room = host_aos ? (int)std::min<int64_t>(max_np, n) : 0;
  if (pack_scratch(d, sizeof(double) * NC * std::max(1, room))) return 1;
  if (launch_pack_tracer(s, cb[0], n, d->origin_dev + 3 * k, d->pack_dev, room, d->count_dev, d->stream, d->fp32)) return 1;
  int cnt = 0;
  NIX_CUDA(cudaMemcpyAsync(&cnt, d->count_dev, sizeof(int), cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  *np = cnt;
  if (!host_aos) return 0;
  if (cnt > max_np) {
    set_error("output buffer too small");
    return 1;
  }
  if (cnt > 0) {
    NIX_CUDA(cudaMemcpyAsync(host_aos, d->pack_dev, sizeof(double) * NC * cnt, cudaMemcpyDeviceToHost, d->stream));
    NIX_CUDA(cudaStreamSynchronize(d->stream));
  }
  return 0;
}

int nixb200_shape_eval(int device, int kind, int order, int n, const double* x, const double* X, double rdx, double dt,
                       double rdt, double* out)
{
  if (!x || !X || !out || n < 0 || order < 1 || order > 4 || kind < 0 || kind > 1) {
    set_error("shape_eval: bad argument (order 1..4, kind 0 = shape_mc, 1 = shape_wt)");
    return 1;
  }
  DeviceGuard dev_guard(device);
  double *dx = nullptr, *dX = nullptr, *dout = nullptr;
  NIX_CUDA(cudaMalloc(&dx, sizeof(double) * std::max(1, n)));
  NIX_CUDA(cudaMalloc(&dX, sizeof(double) * std::max(1, n)));
  NIX_CUDA(cudaMalloc(&dout, sizeof(double) * std::max(1, n) * (order + 1)));
  NIX_CUDA(cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice));
  NIX_CUDA(cudaMemcpy(dX, X, sizeof(double) * n, cudaMemcpyHostToDevice));
  int rc = launch_shape_eval(kind, order, n, dx, dX, rdx, dt, rdt, dout, 0);
  if (!rc) NIX_CUDA(cudaMemcpy(out, dout, sizeof(double) * n * (order + 1), cudaMemcpyDeviceToHost));
  cudaFree(dx);
  cudaFree(dX);
  cudaFree(dout);
  return rc;
}

// Chunk::set_mpi_buffer, chunk.cpp:257-286 (headbyte 0, elembyte 8 * ncomp): host logic, no device needed
int nixb200_halo_layout_dims(const int* dims, int nb, int mode, int* bufsize27, int* bufaddr27)
{
  if (!dims || !bufsize27 || !bufaddr27 || nb < 1 || dims[0] < 1 || dims[1] < 1 || dims[2] < 1) {
    set_error("halo_layout: bad argument");
    return 1;
  }
  if (mode != NIXB200_MODE_FIELD && mode != NIXB200_MODE_CURRENT) {
    set_error("halo_layout: fixed layouts exist for field and current only");
    return 1;
  }
  const int elem = 8 * ((mode == NIXB200_MODE_FIELD) ? 6 : 4);
  int       size = 0;
  for (int s = 0; s < 27; s++) {
    bufaddr27[s] = size;
    if (s == 13) {
      bufsize27[s] = 0;
      continue;
    }
    int e[3] = {s / 9, (s / 3) % 3, s % 3};
    int cnt  = 1;
    for (int a = 0; a < 3; a++) cnt *= (e[a] == 1) ? dims[a] : nb;
    bufsize27[s] = elem * cnt;
    size += bufsize27[s];
  }
  return 0;
}

int nixb200_halo_layout(nixb200_domain* dd, int mode, int* bufsize27, int* bufaddr27)
{
  NIX_ENTER(dd);
  return nixb200_halo_layout_dims(d->geo.N, d->geo.nb, mode, bufsize27, bufaddr27);
}

int nixb200_chunk_halo_pack(nixb200_domain* dd, int k, int mode, void* host_sendbuf)
{
  NIX_ENTER(dd);
  if (check_chunk(d, k) || !host_sendbuf) return 1;
  if (d->fp32) {
    set_error("fp32 mode: the MpiBuffer-layout halo calls are fp64 only");
    return 1;
  }
  int bs[27], ba[27];
  if (nixb200_halo_layout(dd, mode, bs, ba)) return 1;
  size_t total = (size_t)ba[26] + bs[26];
  const double* data = reinterpret_cast<const double*>((mode == NIXB200_MODE_FIELD) ? d->uf : d->uj);
  if (launch_halo_pack(d->geo, k, mode, data, d->halo_buf, d->stream)) return 1;
  NIX_CUDA(cudaMemcpyAsync(host_sendbuf, d->halo_buf, total, cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  return 0;
}

int nixb200_chunk_halo_unpack(nixb200_domain* dd, int k, int mode, const void* host_recvbuf, const int* nbvalid27)
{
  NIX_ENTER(dd);
  if (check_chunk(d, k) || !host_recvbuf) return 1;
  if (d->fp32) {
    set_error("fp32 mode: the MpiBuffer-layout halo calls are fp64 only");
    return 1;
  }
  int bs[27], ba[27];
  if (nixb200_halo_layout(dd, mode, bs, ba)) return 1;
  size_t total = (size_t)ba[26] + bs[26];
  int    valid[27];
  for (int s = 0; s < 27; s++) valid[s] = nbvalid27 ? nbvalid27[s] : 1;
  NIX_CUDA(cudaMemcpyAsync(d->nbvalid_dev, valid, sizeof(valid), cudaMemcpyHostToDevice, d->stream));
  NIX_CUDA(cudaMemcpyAsync(d->halo_buf, host_recvbuf, total, cudaMemcpyHostToDevice, d->stream));
  double* data = reinterpret_cast<double*>((mode == NIXB200_MODE_FIELD) ? d->uf : d->uj);
  if (launch_halo_unpack(d->geo, k, mode, data, d->halo_buf, d->nbvalid_dev, d->stream)) return 1;
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  return 0;
}

int nixb200_domain_reserve(nixb200_domain* dd, int is, int64_t np, int64_t nmove)
{
  NIX_ENTER(dd);
  if (check_species(d, is)) return 1;
  if (np > d->sp[is].cap && grow_particles(d, d->sp[is], np)) return 1;
  if (nmove > d->sp[is].lcap && grow_leavers(d, d->sp[is], nmove, false)) return 1;
  return 0;
}

int nixb200_domain_get_capacity(nixb200_domain* dd, int is, int64_t* np, int64_t* nmove)
{
  NIX_ENTER(dd);
  if (check_species(d, is)) return 1;
  if (np) *np = d->sp[is].cap;
  if (nmove) *nmove = d->sp[is].lcap;
  return 0;
}

int nixb200_domain_get_load(nixb200_domain* dd, double* ms)
{
  NIX_ENTER(dd);
  if (!d || !ms) return 1;
  *ms = 0.0;
  if (!d->timed) return 0;
  NIX_CUDA(cudaEventSynchronize(d->ev1));
  float t = 0.f;
  NIX_CUDA(cudaEventElapsedTime(&t, d->ev0, d->ev1));
  *ms = t;
  return 0;
}

int64_t nixb200_domain_total_particles(nixb200_domain* dd)
{
  Domain* d = D(dd);
  if (!d) return -1;
  DeviceGuard dev_guard(d->desc.device);
  int64_t tot = 0;
  for (auto& s : d->sp) {
    int32_t n = 0;
    if (cudaMemcpyAsync(&n, s.cbase + d->geo.nchunk, sizeof(int32_t), cudaMemcpyDeviceToHost, d->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(d->stream) != cudaSuccess) return -1;
    tot += n;
  }
  return tot;
}

} // extern "C"
