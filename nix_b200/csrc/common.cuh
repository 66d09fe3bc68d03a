// common.cuh -- internal types shared by the sm_100a kernels of libnixb200.so
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/nixb200.h"

namespace nixb200
{
constexpr int NC    = 7; // particle.hpp:18   components per particle (x y z ux uy uz id)
constexpr int LANES = 8; // nix.hpp:105-108   NIX_SIMD_WIDTH: lane count of pcount[Ng+1][8]

// ---------------------------------------------------------------------------------------------
// geometry shared by all chunks of a domain (passed to kernels by value)
// ---------------------------------------------------------------------------------------------
struct Geo {
  int    N[3];    // cells per chunk (z,y,x)
  int    M[3];    // with ghosts
  int    R[3];    // N+1: strides of XtensorParticle::flatindex (xtensor_particle.hpp:231-238)
  int    nc[3];   // cells per axis that can hold particles: N + is_odd
  int    nb, order, is_odd, half;
  int    ncell;   // R[0]*R[1]*R[2]  flat indices in use
  int    nchunk;  // local chunks
  // push work item = a (tile[0] x tile[1] x tile[2]) box of bins of one chunk, one CTA each
  int    tile[3];  // bins per tile (z,y,x)
  int    ntl[3];   // tiles per axis = ceil(nc/tile)
  int    ntile;    // tiles per chunk
  // leaver bookkeeping: a particle that leaves the chunk in direction d can only come from the
  // outermost `slabt` bin layers on the sides where d points outwards (it moves < 1 bin per step),
  // so the per-(bin, direction) leaver counts live in 26 slabs instead of a full [ncell][27] array
  int    slabt;       // slab thickness in bins: 1 for even orders, 2 for odd (bins straddle cells)
  int    slaboff[28]; // first entry of the slab of direction d; slaboff[27] = entries per chunk
  double del[3], rdel[3];
  double glo[3], ghi[3], glen[3]; // global box (set_boundary_periodic, xtensor_particle.hpp:359-369)
  double cc, rc;
};

// per-chunk constants, all precomputed on the host with the reference's expressions so that the
// device never re-derives them with a different rounding (chunk.cpp:210-237,
// xtensor_particle.hpp:332-334)
struct ChunkGeo {
  double lo[3], hi[3]; // chunk range [lo, hi)
  double off[3];       // bin offset       lo - 0.5*del*is_odd
  double hoff[3];      // half-grid offset lo - 0.5*del*(1-is_odd)
  double imin[3];      // position of integer node 0 (cell centre): lo + 0.5*del
  int    nbr[27];      // local index of the neighbour chunk, -1: none, -2-r: on rank r
};

// one species on one device: SoA particle store, double-buffered like xu/xv of the reference
struct SpeciesDev {
  double*  xu;    // [7][cap]
  double*  xv;    // [7][cap]
  int32_t* key;   // [cap]   (chunk*ncell + cell)*8 + lane, -1 = dropped / leaver
  int32_t* ordl;  // [cap]   per-bin member list: position in xu, or 0x80000000|message slot (stable rank)
  int32_t* hist;  // [nchunk*ncell*8]      counts -> consumed as cursors by place
  int32_t* start; // [nchunk*ncell*8 + 1]  exclusive scan of hist (global particle index)
  int32_t* oob;   // [nchunk][8]  row Ng of the reference's pcount
  int32_t* cbase; // [nchunk+1]   first particle of each chunk in xu
  int32_t* cbase_new;
  // migration
  int32_t* slabcnt;  // [nchunk][slaboff[27]] leavers per (slab bin, direction); scanned in place
  int32_t* sendcnt;  // [nchunk][27]
  int32_t* msgoff;   // [nchunk][27]  first message slot of RECEIVE slot e of a chunk (destination-major)
  int32_t* recvoff;  // [nchunk][27]  pre-sort local index of the first particle received in slot e
  int32_t* nleave;   // [1] number of leaver records
  int32_t* nmsg;     // [1] total message particles
  int4*    lrec;     // [lcap] leaver records {i, chunk, slab entry, (rank in bin)<<5 | dir}
  double*  msg;      // [7][lcap] message payload (wrapped positions), SoA
  int32_t* msgkey;   // [lcap]
  // neighbours on other ranks (peer.cu); null / unused for a single-rank domain
  double*  paysend;  // [lcap][7] AoS payload of the leavers bound for other ranks, peer-major
  double*  payrecv;  // [lcap][7] AoS payload received from other ranks, peer-major
  int32_t* ptab;     // per-step tables: spoff[nsend+1] | rpoff[nrecv+1] | rcnt[nrecv]
  double   q, m;
  int64_t  cap, lcap;
};

// one slab crossing a rank boundary: local chunk k, direction / receive slot dir, first cell of the
// slab inside the (peer-major) halo buffer, cells in the slab, and the position of its particle
// count inside the count message of its peer: cidx + species * cstride
struct PeerEntry {
  int k, dir, celloff, cells, cidx, cstride;
};

// tables of the cross-rank exchange, passed to kernels by value; all null for a single-rank domain
struct PeerTabs {
  const PeerEntry* send_ent;     // [nsend] sorted by (peer, sender chunk id, direction)
  const PeerEntry* recv_ent;     // [nrecv] same order as the sender's list
  const int32_t*   send_slot;    // [nchunk][27] -> send entry, -1
  const int32_t*   recv_slot;    // [nchunk][27] -> recv entry, -1
  int              nsend, nrecv;
};

inline __host__ __device__ size_t soa(int comp, size_t cap, size_t i)
{
  return (size_t)comp * cap + i;
}

// ---------------------------------------------------------------------------------------------
// arithmetic helpers.  STRICT = evaluate exactly as written (no FMA contraction) so that push
// results are bit-identical to the reference's scalar templates compiled without contraction.
// ---------------------------------------------------------------------------------------------
// (double: __d*_rn, float: __f*_rn; T is deduced, so mixing a float with a double operand does not compile --
// the fp32 instantiations must not be promoted to fp64 behind our back)
template <bool S, typename T>
__device__ __forceinline__ T mul(T a, T b)
{
  if constexpr (!S) return a * b;
  else if constexpr (sizeof(T) == 8) return __dmul_rn(a, b);
  else return __fmul_rn(a, b);
}
template <bool S, typename T>
__device__ __forceinline__ T add(T a, T b)
{
  if constexpr (!S) return a + b;
  else if constexpr (sizeof(T) == 8) return __dadd_rn(a, b);
  else return __fadd_rn(a, b);
}
template <bool S, typename T>
__device__ __forceinline__ T sub(T a, T b)
{
  if constexpr (!S) return a - b;
  else if constexpr (sizeof(T) == 8) return __dsub_rn(a, b);
  else return __fsub_rn(a, b);
}
// a*b + c
template <bool S, typename T>
__device__ __forceinline__ T mad(T a, T b, T c)
{
  if constexpr (!S) {
    if constexpr (sizeof(T) == 8) return fma(a, b, c);
    else return fmaf(a, b, c);
  } else return add<true>(mul<true>(a, b), c);
}
template <bool S, typename T>
__device__ __forceinline__ T div_(T a, T b)
{
  if constexpr (!S) return a / b;
  else if constexpr (sizeof(T) == 8) return __ddiv_rn(a, b);
  else return __fdiv_rn(a, b);
}
template <bool S, typename T>
__device__ __forceinline__ T sqrt_(T a)
{
  if constexpr (sizeof(T) == 8) return S ? __dsqrt_rn(a) : sqrt(a);
  else return S ? __fsqrt_rn(a) : sqrtf(a);
}

// primitives.hpp:46-58 -- always exact (counts and permutations must be bit-exact)
__device__ __forceinline__ int digitize(double x, double xmin, double rdx)
{
  return (int)floor(__dmul_rn(__dsub_rn(x, xmin), rdx));
}
__device__ __forceinline__ int digitize(float x, float xmin, float rdx)
{
  return (int)floorf(__fmul_rn(__fsub_rn(x, xmin), rdx));
}

// The real type of a domain: fp64 is the reference's (nix.hpp:76-78 has real = float64 only); fp32 is the
// north star's second mode.  NCT = words of T per particle in the SoA store: x y z ux uy uz + the 64-bit id.
template <typename T>
struct Real;
template <>
struct Real<double> {
  using vec2 = double2;
  static constexpr int NCT = 7;
  static __device__ __forceinline__ double2 make2(double a, double b) { return make_double2(a, b); }
};
template <>
struct Real<float> {
  using vec2 = float2;
  static constexpr int NCT = 8; // the id occupies two words
  static __device__ __forceinline__ float2 make2(float a, float b) { return make_float2(a, b); }
};

// words per cell of the E/B array on the device: the reference's 6 doubles (48 B), or 8 floats (32 B, two
// of them padding) because TMA strides must be multiples of 16 bytes
template <typename T>
__host__ __device__ constexpr int field_stride()
{
  return sizeof(T) == 8 ? 6 : 8;
}

// per-chunk constants in the real type of the kernel (staged in shared memory from ChunkGeo)
template <typename T>
struct ChunkGeoT {
  T   lo[3], hi[3], off[3], hoff[3], imin[3];
  int nbr[27];
};
template <typename T>
__device__ __forceinline__ void stage_chunk_geo(ChunkGeoT<T>* dst, const ChunkGeo* src, int tid, int nthreads)
{
  const double* sd = reinterpret_cast<const double*>(src);
  for (int t = tid; t < 15; t += nthreads) reinterpret_cast<T*>(dst)[t] = (T)sd[t];
  for (int t = tid; t < 27; t += nthreads) dst->nbr[t] = src->nbr[t];
}

// entry of bin (b[0],b[1],b[2]) in the slab of direction d = 9*ez+3*ey+ex (0/1/2 = -/centre/+);
// -1 when the bin is not inside that slab.  Entries of one slab follow the flat bin order, i.e.
// the particle order of the cell-sorted container.
__host__ __device__ inline int slab_entry(const Geo& g, int d, const int* b)
{
  int e[3] = {d / 9, (d / 3) % 3, d % 3};
  int pos = 0;
  for (int a = 0; a < 3; a++) {
    int lo = (e[a] == 2) ? g.nc[a] - g.slabt : 0;
    int n  = (e[a] == 1) ? g.nc[a] : g.slabt;
    int r  = b[a] - lo;
    if (r < 0 || r >= n) return -1;
    pos = pos * n + r;
  }
  return g.slaboff[d] + pos;
}

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void        set_error(const std::string& msg);
extern std::atomic<int64_t> g_launches;

#define NIX_CUDA(call)                                                                           \
  do {                                                                                           \
    cudaError_t err__ = (call);                                                                  \
    if (err__ != cudaSuccess) {                                                                  \
      nixb200::set_error(std::string(#call) + ": " + cudaGetErrorString(err__) + " (" +          \
                         __FILE__ + ":" + std::to_string(__LINE__) + ")");                       \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)

#define NIX_LAUNCHED()                                                                           \
  do {                                                                                           \
    nixb200::g_launches++;                                                                       \
    NIX_CUDA(cudaGetLastError());                                                                \
  } while (0)

// ---------------------------------------------------------------------------------------------
// kernel launchers (one per .cu file); all enqueue on `st` and return 0 / non-zero
// ---------------------------------------------------------------------------------------------
struct PushArgs {
  Geo              geo;
  const ChunkGeo*  cg;
  const void*      uf; // [nchunk][Mz][My][Mx][6]  (real type of the domain)
  void*            uj; // [nchunk][Mz][My][Mx][4]
  bool             fp32 = false;
  SpeciesDev       sp;
  double           delt;
  int*             err;    // err[0]: NIXB200_ERR_* bits; err[1]: sticky "store overflowed" flag (kernels exit)
  int              pusher; // NIXB200_PUSH_*
};

// ev: null, or four events recorded around the two kernels: ev[0] k_push ev[1] | ev[2] k_deposit ev[3]
int launch_push_deposit(const PushArgs& a, const CUtensorMap* tmap, bool strict, cudaStream_t st,
                        cudaEvent_t* ev = nullptr);
int push_deposit_prepare(int order, bool fp32); // per-device kernel attributes (call with the device current)
size_t push_smem_bytes(const Geo& g);
void   push_tile_box(int order, int& ez, int& ey, int& ex); // nodes per axis of the staged E/B tile
int    choose_push_tile(Geo& g); // fills tile / ntl / ntile; non-zero if nothing fits

int launch_count_only(const Geo& g, const ChunkGeo* cg, SpeciesDev& sp, int* err, cudaStream_t st, bool fp32);
int launch_mig_scan(const Geo& g, SpeciesDev& sp, cudaStream_t st);
int launch_peer_counts(const Geo& g, const SpeciesDev& sp, const PeerTabs& pt, int is, int32_t* cnt_send,
                       cudaStream_t st);
int launch_mig_route(const Geo& g, const ChunkGeo* cg, SpeciesDev& sp, const PeerTabs& pt, int* err,
                     cudaStream_t st, bool fp32);
int launch_stats(const Geo& g, const SpeciesDev& sp, int32_t* out4, cudaStream_t st);
int launch_mig_recv(const Geo& g, const ChunkGeo* cg, SpeciesDev& sp, const PeerTabs& pt, int nrecv_particles,
                    int* err, cudaStream_t st, bool fp32);
int launch_sort(const Geo& g, const ChunkGeo* cg, SpeciesDev& sp, int* err, void* scan_tmp,
                cudaStream_t st, bool fp32);
size_t scan_tmp_bytes(size_t n);
int launch_hist_from_start(const int32_t* start, int32_t* hist, size_t n, cudaStream_t st);
int launch_scan_only(const Geo& g, SpeciesDev& sp, int* err, void* scan_tmp, cudaStream_t st);

int launch_halo_field(const Geo& g, const ChunkGeo* cg, void* uf, const PeerTabs& pt, const void* recvbuf,
                      cudaStream_t st, bool fp32);
int launch_halo_current(const Geo& g, const ChunkGeo* cg, void* uj, const PeerTabs& pt, const void* recvbuf,
                        cudaStream_t st, bool fp32);
int launch_peer_pack(const Geo& g, int mode, const void* data, const PeerTabs& pt, void* sendbuf,
                     cudaStream_t st, bool fp32, int ncomp_moment = 0);
int launch_halo_moment(const Geo& g, const ChunkGeo* cg, void* um, int ncomp, const PeerTabs& pt, const void* recvbuf,
                       cudaStream_t st, bool fp32);

// diagnostics / output path (diag.cu)
int pack_count(const Geo& g, int decimate, int nc);
int launch_pack_grid(const Geo& g, const void* chunk_ptr, bool colocate, int decimate, int nc, int fc, double* out,
                     cudaStream_t st, bool fp32);
int launch_pack_tracer(const SpeciesDev& sp, int first, int np, const double* origin3, double* out, int max_out,
                       int* count_dev, cudaStream_t st, bool fp32);
int launch_moment(const Geo& g, const ChunkGeo* cg, const SpeciesDev& sp, int is, int ns, void* um, cudaStream_t st,
                  bool fp32);
int launch_shape_eval(int kind, int order, int n, const double* x, const double* X, double rdx, double dt, double rdt,
                      double* out, cudaStream_t st);
int launch_halo_pack(const Geo& g, int k, int mode, const double* data, double* buf, cudaStream_t st);
int launch_halo_unpack(const Geo& g, int k, int mode, double* data, const double* buf,
                       const int* nbvalid_dev, cudaStream_t st);

// Yee FDTD on the device-resident grid (fdtd.cu)
int launch_push_bfd(const Geo& g, void* uf, double delt, int ext, cudaStream_t st, bool fp32);
int launch_push_efd(const Geo& g, void* uf, const void* uj, double delt, double cfj, cudaStream_t st, bool fp32);
int launch_field_energy(const Geo& g, const void* uf, double* out, cudaStream_t st, bool fp32);

// interior cells of all chunks <-> dense [chunk][Nz][Ny][Nx][ncomp] (pack: full -> dense)
int launch_interior(bool pack, double* full, double* dense, const Geo& g, int ncomp, cudaStream_t st);
// fp32 mode: conversions at the (fp64) C-ABI boundary
int launch_aos_to_soa_f32(const double* aos, float* soa_base, size_t cap, size_t first, size_t n, const int32_t* cbase,
                          int nchunk, const double* origin, const double* extent /* z,y,x */, cudaStream_t st);
int launch_soa_to_aos_f32(const float* soa_base, double* aos, size_t cap, size_t first, size_t n, const double* origin3,
                          cudaStream_t st);
int launch_cells_convert(bool to_dev, double* host_layout, float* dev_layout, size_t ncell, int nc, int fc, cudaStream_t st);
int launch_np_from_cbase(const int32_t* cbase, int nch, int64_t* np, cudaStream_t st);
int launch_interior_f32(bool pack, float* full, double* dense, const Geo& g, int ncomp, int fc, cudaStream_t st);
int launch_aos_to_soa(const double* aos, double* soa_base, size_t cap, size_t first, size_t n,
                      cudaStream_t st);
int launch_soa_to_aos(const double* soa_base, double* aos, size_t cap, size_t first, size_t n,
                      cudaStream_t st);
} // namespace nixb200
