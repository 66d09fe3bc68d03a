// sort.cu -- per-cell particle count / sort and particle migration between chunks (sm_100a)
//
// Replaces, for every chunk of a device at once:
//   XtensorParticle::count   xtensor_particle.hpp:324-357
//   XtensorParticle::sort    xtensor_particle.hpp:260-321
//   XtensorHaloParticle3D    xtensor_halo3d.hpp:251-557 (classify, pack, append, wrap, count, sort)
//
// The reference sorts each chunk with a serial counting sort whose net effect is a STABLE sort of
// the chunk's pre-sort array by the key  cell*8 + (ip % 8)  (ip = index in the pre-sort array,
// received particles appended behind the residents in direction order), dropping out-of-bounds
// particles.  Here ONE global key space covers every chunk of the device:
//     key = (chunk*ncell + cell)*8 + lane,
// the histogram over it is the reference's pcount, its exclusive scan gives pindex and the new
// chunk bases, and stability is obtained without any ordered atomics: `place` drops each
// particle's pre-sort position `ord` into the slot range of its bin in arbitrary order; `gather`
// then runs over the DESTINATION slots: slot j holds some member of its bin, ranks it by counting
// the smaller `ord`s of the (tiny) bin and moves the particle to start[key] + rank.  A warp thus
// writes a permutation of its own 32 consecutive slots (full sectors, no write amplification) and
// reads the pre-sort neighbourhood of the same cells.  All integer work -> bit-exact.
#include "common.cuh"

namespace nixb200
{
namespace
{
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS   = 16;
constexpr int SCAN_TILE    = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int find_chunk(const int32_t* __restrict__ cbase, int nchunk, int i)
{
  int lo = 0, hi = nchunk; // cbase[lo] <= i < cbase[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (cbase[mid] <= i) lo = mid;
    else hi = mid;
  }
  return lo;
}

// direction code 9*dz+3*dy+dx of a position relative to the chunk (xtensor_halo3d.hpp:291-293)
template <typename T>
__device__ __forceinline__ int dir_code(const ChunkGeo& c, T x, T y, T z)
{
  int dz = (z >= (T)c.hi[0]) - (z < (T)c.lo[0]) + 1;
  int dy = (y >= (T)c.hi[1]) - (y < (T)c.lo[1]) + 1;
  int dx = (x >= (T)c.hi[2]) - (x < (T)c.lo[2]) + 1;
  return 9 * dz + 3 * dy + dx;
}

// flat cell index of an in-bounds position (xtensor_particle.hpp:345-348)
template <typename T>
__device__ __forceinline__ int cell_index(const Geo& g, const ChunkGeo& c, T x, T y, T z)
{
  int ix = digitize(x, (T)c.off[2], (T)g.rdel[2]);
  int iy = digitize(y, (T)c.off[1], (T)g.rdel[1]);
  int iz = digitize(z, (T)c.off[0], (T)g.rdel[0]);
  return (iz * g.R[1] + iy) * g.R[2] + ix;
}

// ---------------------------------------------------------------------------------------------
// count only (initial binning): every particle is a resident or is dropped
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_count(Geo g, const ChunkGeo* __restrict__ cg, SpeciesDev sp)
{
  const int ntot = sp.cbase[g.nchunk];
  const T* __restrict__ xu = reinterpret_cast<const T*>(sp.xu);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ntot; i += gridDim.x * blockDim.x) {
    int             ch = find_chunk(sp.cbase, g.nchunk, i);
    const ChunkGeo& c  = cg[ch];
    T               x  = xu[soa(0, sp.cap, i)];
    T               y  = xu[soa(1, sp.cap, i)];
    T               z  = xu[soa(2, sp.cap, i)];
    int             lane = (i - sp.cbase[ch]) & (LANES - 1);
    int             key  = -1;
    if (dir_code(c, x, y, z) == 13) {
      key = (ch * g.ncell + cell_index(g, c, x, y, z)) * LANES + lane;
      atomicAdd(&sp.hist[key], 1);
    } else {
      atomicAdd(&sp.oob[ch * LANES + lane], 1);
    }
    sp.key[i] = key;
  }
}

// ---------------------------------------------------------------------------------------------
// migration bookkeeping
// ---------------------------------------------------------------------------------------------
// exclusive scan of the per-bin leaver counts inside the slab of each (chunk, direction); slab
// entries follow the flat bin order = particle order.  One block per (chunk, direction).
__global__ void __launch_bounds__(128) k_mig_scan(Geo g, SpeciesDev sp)
{
  __shared__ int s_part[128];
  const int ch  = blockIdx.x / 27;
  const int dir = blockIdx.x % 27;
  if (dir == 13) {
    if (threadIdx.x == 0) sp.sendcnt[ch * 27 + dir] = 0;
    return;
  }
  int32_t*  sc  = sp.slabcnt + (size_t)ch * g.slaboff[27] + g.slaboff[dir];
  const int n   = g.slaboff[dir + 1] - g.slaboff[dir];
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int b = threadIdx.x * per, e = min(n, b + per);
  int       sum = 0;
  for (int f = b; f < e; f++) sum += sc[f];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int t = 0; t < blockDim.x; t++) {
      int v     = s_part[t];
      s_part[t] = run;
      run += v;
    }
    sp.sendcnt[ch * 27 + dir] = run;
  }
  __syncthreads();
  int run = s_part[threadIdx.x];
  for (int f = b; f < e; f++) {
    int v = sc[f];
    sc[f] = run;
    run += v;
  }
}

// message slots and append offsets of every receive slot (pre_unpack/unpack order: slots in
// (iz,iy,ix) order, xtensor_halo3d.hpp:440-461,507-524).  The message buffer is DESTINATION-major:
// the particles chunk B receives sit in the order B appends them, so that the order of message
// slots inside a chunk equals the order of the pre-sort indices (what the stable rank needs).
__global__ void k_mig_offsets(Geo g, const ChunkGeo* __restrict__ cg, SpeciesDev sp, PeerTabs pt)
{
  // received from another rank: count table of this step, rcnt[] behind spoff[nsend+1] | rpoff[nrecv+1]
  const int32_t* rcnt = (pt.recv_slot != nullptr) ? sp.ptab + (pt.nsend + 1) + (pt.nrecv + 1) : nullptr;
  __shared__ int s_part[1024];
  const int      tid = threadIdx.x;
  const int      per = (g.nchunk + blockDim.x - 1) / blockDim.x;
  const int      b = tid * per, e = min(g.nchunk, b + per);
  auto recv_cnt = [&](int ch, int s) {
    int nb = cg[ch].nbr[s];
    if (s == 13) return 0;
    if (nb >= 0) return sp.sendcnt[nb * 27 + (26 - s)];
    if (rcnt != nullptr) {
      int j = pt.recv_slot[ch * 27 + s];
      if (j >= 0) return rcnt[j];
    }
    return 0;
  };
  int sum = 0;
  for (int ch = b; ch < e; ch++)
    for (int s = 0; s < 27; s++) sum += recv_cnt(ch, s);
  s_part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < blockDim.x; t++) {
      int v     = s_part[t];
      s_part[t] = run;
      run += v;
    }
    *sp.nmsg = run;
  }
  __syncthreads();
  int run = s_part[tid];
  for (int ch = b; ch < e; ch++) {
    int running = sp.cbase[ch + 1] - sp.cbase[ch];
    for (int s = 0; s < 27; s++) {
      int cnt                 = recv_cnt(ch, s);
      sp.msgoff[ch * 27 + s]  = run;     // first message slot of receive slot s of chunk ch
      sp.recvoff[ch * 27 + s] = running; // pre-sort local index of its first particle
      run += cnt;
      running += cnt;
    }
  }
}

// a migrating particle arrives in chunk B as its pre-sort particle ipB: periodic wrap, bin, count,
// message slot m (post_unpack: set_boundary_periodic + count(reset=false), xtensor_halo3d.hpp:541-548).
// dir = the direction it left its old chunk in.  fp64: positions are global and wrap around the box;
// fp32: positions are relative to the chunk's origin (an fp32 global coordinate would lose 1e-5 of a cell at
// x ~ 100 cells), so arriving means shifting by one chunk extent and periodicity is in the neighbour table.
template <typename T>
__device__ __forceinline__ void deliver(const Geo& g, const ChunkGeo* __restrict__ cg, const SpeciesDev& sp, int B,
                                        int m, int ipB, T* v, int dir)
{
  constexpr int NCT = Real<T>::NCT;
  if constexpr (sizeof(T) == 8) {
    // xtensor_particle.hpp:371-375   x += (x < X1)*L - (x >= X2)*L   (z,y,x stored as v[2],v[1],v[0])
#pragma unroll
    for (int a = 0; a < 3; a++) {
      double p  = v[2 - a];
      double L  = g.glen[a];
      double s1 = (p < g.glo[a]) ? L : 0.0;
      double s2 = (p >= g.ghi[a]) ? L : 0.0;
      v[2 - a]  = __dadd_rn(p, __dsub_rn(s1, s2));
    }
  } else {
    const int e[3] = {dir / 9 - 1, (dir / 3) % 3 - 1, dir % 3 - 1};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const T ext = (T)(g.N[a] * g.del[a]);
      T       p   = v[2 - a] - (T)e[a] * ext; // e = +1: exact (Sterbenz); e = -1: rounds
      // a particle that left through the low face by less than half an ulp of the extent must not land ON the
      // neighbour's upper bound (it would be counted out of bounds and dropped, 1e-7-rare in fp32)
      if (e[a] < 0 && p >= ext) p = __int_as_float(__float_as_int(ext) - 1);
      v[2 - a] = p;
    }
  }
  const ChunkGeo& c    = cg[B];
  int             lane = ipB & (LANES - 1);
  int             key  = -1;
  if (dir_code<T>(c, v[0], v[1], v[2]) == 13) {
    key = (B * g.ncell + cell_index<T>(g, c, v[0], v[1], v[2])) * LANES + lane;
    atomicAdd(&sp.hist[key], 1);
  } else {
    atomicAdd(&sp.oob[B * LANES + lane], 1);
  }
  T* msg = reinterpret_cast<T*>(sp.msg);
#pragma unroll
  for (int k = 0; k < NCT; k++) msg[soa(k, sp.lcap, m)] = v[k];
  sp.msgkey[m] = key;
}

// The leaver list or the message buffer overflowed (k_push dropped records, or more particles arrive
// than fit): message slots that nobody fills must not be consumed with last step's content.  Normally
// exits at once.
__global__ void k_mig_guard(SpeciesDev sp, int* err)
{
  if (*sp.nleave <= sp.lcap && *sp.nmsg <= sp.lcap) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(err, NIXB200_ERR_CAPACITY);
  for (int64_t m = blockIdx.x * blockDim.x + threadIdx.x; m < sp.lcap; m += (int64_t)gridDim.x * blockDim.x)
    sp.msgkey[m] = -1;
}

// what the host needs to grow the stores in time: particles after the sort, leaver records, messages
__global__ void k_stats(Geo g, SpeciesDev sp, int32_t* __restrict__ out4)
{
  const size_t nkey = (size_t)g.nchunk * g.ncell * LANES;
  out4[0] = sp.start[nkey];
  out4[1] = *sp.nleave;
  out4[2] = *sp.nmsg;
  out4[3] = 0;
}

// one thread per leaver: destination chunk, pre-sort index there, periodic wrap, count
// (post_unpack: set_boundary_periodic + count(reset=false), xtensor_halo3d.hpp:541-548)
template <typename T>
__global__ void k_mig_key(Geo g, const ChunkGeo* __restrict__ cg, SpeciesDev sp, PeerTabs pt, int* err)
{
  constexpr int NCT = Real<T>::NCT;
  const T* __restrict__ xu = reinterpret_cast<const T*>(sp.xu);
  T* paysend = reinterpret_cast<T*>(sp.paysend);
  const int nl = min(*sp.nleave, (int)sp.lcap);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nl; j += gridDim.x * blockDim.x) {
    int4 r    = sp.lrec[j];
    int  i    = r.x;
    int  A    = r.y;
    int  dir  = r.w & 31;
    int  B    = cg[A].nbr[dir];
    int idx   = sp.slabcnt[(size_t)A * g.slaboff[27] + r.z] + (r.w >> 5);
    if (B < 0) {
      // neighbour on another rank: raw payload (the receiver wraps / shifts and bins it) at the slot the
      // reference's pack order gives it, peer-major
      int js = (pt.send_slot != nullptr) ? pt.send_slot[A * 27 + dir] : -1;
      if (js < 0) continue; // no neighbour
      size_t pos = (size_t)sp.ptab[js] + idx;
      if (pos >= (size_t)sp.lcap) {
        atomicOr(err, NIXB200_ERR_CAPACITY);
        continue;
      }
#pragma unroll
      for (int k = 0; k < NCT; k++) paysend[pos * NCT + k] = xu[soa(k, sp.cap, i)];
      continue;
    }
    int m     = sp.msgoff[B * 27 + (26 - dir)] + idx;
    int ipB   = sp.recvoff[B * 27 + (26 - dir)] + idx;
    if (m >= sp.lcap) {
      atomicOr(err, NIXB200_ERR_CAPACITY);
      continue;
    }
    T v[NCT];
#pragma unroll
    for (int k = 0; k < NCT; k++) v[k] = xu[soa(k, sp.cap, i)];
    deliver<T>(g, cg, sp, B, m, ipB, v, dir);
  }
}

// particles received from other ranks: one thread per particle of the peer-major payload
template <typename T>
__global__ void k_mig_recv(Geo g, const ChunkGeo* __restrict__ cg, SpeciesDev sp, PeerTabs pt, int ntot, int* err)
{
  constexpr int NCT = Real<T>::NCT;
  const T* __restrict__ payrecv = reinterpret_cast<const T*>(sp.payrecv);
  const int32_t* rpoff = sp.ptab + (pt.nsend + 1);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntot; t += gridDim.x * blockDim.x) {
    int lo = 0, hi = pt.nrecv; // rpoff[lo] <= t < rpoff[hi]
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (rpoff[mid] <= t) lo = mid;
      else hi = mid;
    }
    const PeerEntry en  = pt.recv_ent[lo];
    const int       idx = t - rpoff[lo];
    const int       m   = sp.msgoff[en.k * 27 + en.dir] + idx;
    const int       ipB = sp.recvoff[en.k * 27 + en.dir] + idx;
    if (m >= sp.lcap) {
      atomicOr(err, NIXB200_ERR_CAPACITY);
      continue;
    }
    T v[NCT];
#pragma unroll
    for (int k = 0; k < NCT; k++) v[k] = payrecv[(size_t)t * NCT + k];
    deliver<T>(g, cg, sp, en.k, m, ipB, v, 26 - en.dir); // received in slot e = sent in direction 26 - e
  }
}

// particle counts of the slabs bound for other ranks -> the count message of each peer
__global__ void k_peer_counts(SpeciesDev sp, PeerTabs pt, int is, int32_t* __restrict__ out)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= pt.nsend) return;
  const PeerEntry en = pt.send_ent[j];
  out[en.cidx + is * en.cstride] = sp.sendcnt[en.k * 27 + en.dir];
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of the histogram (reduce / scan block sums / apply)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v)
{
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan of one value per thread over the block; returns the block total in `total`
__device__ __forceinline__ int block_excl_scan(int v, int& total)
{
  __shared__ int s_w[32];
  const int      lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int            inc = warp_incl_scan(v);
  if (lane == 31) s_w[w] = inc;
  __syncthreads();
  if (w == 0) {
    int t = (lane < nw) ? s_w[lane] : 0;
    t     = warp_incl_scan(t);
    s_w[lane] = t;
  }
  __syncthreads();
  int base = (w > 0) ? s_w[w - 1] : 0;
  total    = s_w[nw - 1];
  __syncthreads();
  return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int32_t* __restrict__ in, size_t n,
                                                              int32_t* __restrict__ bsum)
{
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  int    s    = 0;
  if (base + SCAN_ITEMS <= n) {
    const int4* p = reinterpret_cast<const int4*>(in + base);
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS / 4; q++) {
      int4 v = p[q];
      s += v.x + v.y + v.z + v.w;
    }
  } else {
    for (int q = 0; q < SCAN_ITEMS; q++)
      if (base + q < n) s += in[base + q];
  }
  int total;
  block_excl_scan(s, total);
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_bsum(int32_t* bsum, int nblk, int32_t* total_out)
{
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nblk; b0 += blockDim.x) {
    int i = b0 + threadIdx.x;
    int v = (i < nblk) ? bsum[i] : 0;
    int total;
    int ex    = block_excl_scan(v, total);
    int carry = s_carry;
    if (i < nblk) bsum[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = s_carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const int32_t* __restrict__ in, size_t n,
                                                             const int32_t* __restrict__ bsum,
                                                             int32_t* __restrict__ out)
{
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  int    v[SCAN_ITEMS];
  int    s = 0;
  if (base + SCAN_ITEMS <= n) {
    const int4* p = reinterpret_cast<const int4*>(in + base);
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS / 4; q++) {
      int4 t       = p[q];
      v[4 * q + 0] = t.x;
      v[4 * q + 1] = t.y;
      v[4 * q + 2] = t.z;
      v[4 * q + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++) v[q] = (base + q < n) ? in[base + q] : 0;
  }
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++) s += v[q];
  int total;
  int run = bsum[blockIdx.x] + block_excl_scan(s, total);
  if (base + SCAN_ITEMS <= n) {
    int4* o = reinterpret_cast<int4*>(out + base);
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS / 4; q++) {
      int4 t;
      t.x = run;
      run += v[4 * q + 0];
      t.y = run;
      run += v[4 * q + 1];
      t.z = run;
      run += v[4 * q + 2];
      t.w = run;
      run += v[4 * q + 3];
      o[q] = t;
    }
  } else {
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++) {
      if (base + q < n) out[base + q] = run;
      run += v[q];
    }
  }
}

// new chunk bases = scan value at the first key of every chunk; start[n] = total
__global__ void k_chunk_bases(Geo g, SpeciesDev sp, const int32_t* __restrict__ total, int* err)
{
  const size_t n = (size_t)g.nchunk * g.ncell * LANES;
  int          c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0) {
    sp.start[n] = *total;
    if (*total > sp.cap) { // the sorted array would not fit: place / gather and every later kernel stand down
      atomicOr(err, NIXB200_ERR_CAPACITY);
      err[1] = 1;
    }
  }
  if (c < g.nchunk) sp.cbase_new[c] = sp.start[(size_t)c * g.ncell * LANES];
  if (c == g.nchunk) sp.cbase_new[c] = *total;
}

// ---------------------------------------------------------------------------------------------
// place: drop each particle's pre-sort index into its bin (any order)
// ---------------------------------------------------------------------------------------------
__global__ void k_place(Geo g, SpeciesDev sp, const int* __restrict__ err)
{
  if (err[1]) return;
  const int ntot = sp.cbase[g.nchunk];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ntot; i += gridDim.x * blockDim.x) {
    int k = sp.key[i];
    if (k < 0) continue;
    int slot      = sp.start[k] + atomicSub(&sp.hist[k], 1) - 1;
    sp.ordl[slot] = i; // residents: position in xu (same order as the pre-sort index inside a chunk)
  }
}

__global__ void k_place_msg(Geo g, SpeciesDev sp, const int* __restrict__ err)
{
  if (err[1]) return;
  const int nm = min(*sp.nmsg, (int)sp.lcap);
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < nm; m += gridDim.x * blockDim.x) {
    int k = sp.msgkey[m];
    if (k < 0) continue;
    int slot      = sp.start[k] + atomicSub(&sp.hist[k], 1) - 1;
    sp.ordl[slot] = (int)(0x80000000u | (unsigned)m); // received: behind every resident, in append order
  }
}

// ---------------------------------------------------------------------------------------------
// gather: for every destination slot j of the sorted array: the member v = ordl[j] of the bin that
// owns j goes to start[key] + (number of members with a smaller pre-sort position)
// (xtensor_particle.hpp:303-313).  Unsigned compare: received particles (top bit set) follow the
// residents.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_gather(Geo g, SpeciesDev sp, const int* __restrict__ err)
{
  if (err[1]) return;
  constexpr int NCT = Real<T>::NCT;
  const size_t nkey = (size_t)g.nchunk * g.ncell * LANES;
  const int    ntot = min(sp.start[nkey], (int)sp.cap);
  const int32_t* __restrict__ ordl = sp.ordl;
  const T* __restrict__ xu  = reinterpret_cast<const T*>(sp.xu);
  const T* __restrict__ msg = reinterpret_cast<const T*>(sp.msg);
  T* __restrict__ xv        = reinterpret_cast<T*>(sp.xv);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ntot; j += gridDim.x * blockDim.x) {
    const int      v   = ordl[j];
    const bool     res = v >= 0;
    const int      m   = v & 0x7fffffff;
    const int      k   = res ? sp.key[m] : sp.msgkey[m];
    const int      lo = sp.start[k], hi = sp.start[k + 1];
    const unsigned uv = (unsigned)v;
    int            r  = 0;
    for (int e = lo; e < hi; e++) r += ((unsigned)ordl[e] < uv) ? 1 : 0;
    const int    dst = lo + r;
    const T*     src = res ? xu : msg;
    const size_t scp = res ? (size_t)sp.cap : (size_t)sp.lcap;
    T            val[NCT];
#pragma unroll
    for (int c = 0; c < NCT; c++) val[c] = src[soa(c, scp, m)];
#pragma unroll
    for (int c = 0; c < NCT; c++) xv[soa(c, sp.cap, dst)] = val[c];
  }
}

inline int grid_for(size_t n, int threads, int max_blocks = 148 * 16)
{
  size_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > (size_t)max_blocks) b = max_blocks;
  return (int)b;
}
} // namespace

size_t scan_tmp_bytes(size_t n)
{
  size_t nblk = (n + SCAN_TILE - 1) / SCAN_TILE;
  return (nblk + 2) * sizeof(int32_t);
}

int launch_count_only(const Geo& g, const ChunkGeo* cg, SpeciesDev& sp, int* err, cudaStream_t st, bool fp32)
{
  (void)err;
  NIX_CUDA(cudaMemsetAsync(sp.oob, 0, sizeof(int32_t) * g.nchunk * LANES, st));
  NIX_CUDA(cudaMemsetAsync(sp.nleave, 0, sizeof(int32_t), st));
  NIX_CUDA(cudaMemsetAsync(sp.nmsg, 0, sizeof(int32_t), st));
  if (fp32) k_count<float><<<grid_for(sp.cap, 256), 256, 0, st>>>(g, cg, sp);
  else k_count<double><<<grid_for(sp.cap, 256), 256, 0, st>>>(g, cg, sp);
  NIX_LAUNCHED();
  return 0;
}

// after push_deposit (which filled key/hist for residents and the leaver records)
int launch_mig_scan(const Geo& g, SpeciesDev& sp, cudaStream_t st)
{
  k_mig_scan<<<g.nchunk * 27, 128, 0, st>>>(g, sp);
  NIX_LAUNCHED();
  return 0;
}

int launch_peer_counts(const Geo& g, const SpeciesDev& sp, const PeerTabs& pt, int is, int32_t* cnt_send,
                       cudaStream_t st)
{
  (void)g;
  if (pt.nsend == 0) return 0;
  k_peer_counts<<<(pt.nsend + 127) / 128, 128, 0, st>>>(sp, pt, is, cnt_send);
  NIX_LAUNCHED();
  return 0;
}

int launch_mig_route(const Geo& g, const ChunkGeo* cg, SpeciesDev& sp, const PeerTabs& pt, int* err,
                     cudaStream_t st, bool fp32)
{
  k_mig_offsets<<<1, 1024, 0, st>>>(g, cg, sp, pt);
  NIX_LAUNCHED();
  k_mig_guard<<<148, 256, 0, st>>>(sp, err);
  NIX_LAUNCHED();
  if (fp32) k_mig_key<float><<<grid_for(sp.lcap, 256, 148 * 4), 256, 0, st>>>(g, cg, sp, pt, err);
  else k_mig_key<double><<<grid_for(sp.lcap, 256, 148 * 4), 256, 0, st>>>(g, cg, sp, pt, err);
  NIX_LAUNCHED();
  return 0;
}

int launch_stats(const Geo& g, const SpeciesDev& sp, int32_t* out4, cudaStream_t st)
{
  k_stats<<<1, 1, 0, st>>>(g, sp, out4);
  NIX_LAUNCHED();
  return 0;
}

int launch_mig_recv(const Geo& g, const ChunkGeo* cg, SpeciesDev& sp, const PeerTabs& pt, int nrecv_particles,
                    int* err, cudaStream_t st, bool fp32)
{
  if (nrecv_particles <= 0) return 0;
  if (fp32) k_mig_recv<float><<<grid_for(nrecv_particles, 256, 148 * 4), 256, 0, st>>>(g, cg, sp, pt, nrecv_particles, err);
  else k_mig_recv<double><<<grid_for(nrecv_particles, 256, 148 * 4), 256, 0, st>>>(g, cg, sp, pt, nrecv_particles, err);
  NIX_LAUNCHED();
  return 0;
}

// counts of a key range recovered from a scan (rebalance.cu: the chunks that stay keep their per-(cell, lane)
// counts; re-sorting them would regroup the lanes, the reference's key being cell * 8 + position % 8)
__global__ void k_hist_from_start(const int32_t* __restrict__ start, int32_t* __restrict__ hist, size_t n)
{
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
    hist[j] = start[j + 1] - start[j];
}

int launch_hist_from_start(const int32_t* start, int32_t* hist, size_t n, cudaStream_t st)
{
  if (n == 0) return 0;
  k_hist_from_start<<<grid_for(n, 256), 256, 0, st>>>(start, hist, n);
  NIX_LAUNCHED();
  return 0;
}

// hist -> start and the chunk bases only (no particle moves); hist is cleared afterwards, as a sort leaves it
int launch_scan_only(const Geo& g, SpeciesDev& sp, int* err, void* scan_tmp, cudaStream_t st)
{
  const size_t n     = (size_t)g.nchunk * g.ncell * LANES;
  const int    nblk  = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  int32_t*     bsum  = reinterpret_cast<int32_t*>(scan_tmp);
  int32_t*     total = bsum + nblk;
  k_scan_reduce<<<nblk, SCAN_THREADS, 0, st>>>(sp.hist, n, bsum);
  NIX_LAUNCHED();
  k_scan_bsum<<<1, 1024, 0, st>>>(bsum, nblk, total);
  NIX_LAUNCHED();
  k_scan_apply<<<nblk, SCAN_THREADS, 0, st>>>(sp.hist, n, bsum, sp.start);
  NIX_LAUNCHED();
  k_chunk_bases<<<(g.nchunk + 1 + 255) / 256, 256, 0, st>>>(g, sp, total, err);
  NIX_LAUNCHED();
  NIX_CUDA(cudaMemsetAsync(sp.hist, 0, sizeof(int32_t) * n, st));
  return 0;
}

// hist -> start, place, scatter; leaves the sorted particles in sp.xv and the new chunk bases in
// sp.cbase_new (the caller swaps, xtensor_particle.hpp:317)
int launch_sort(const Geo& g, const ChunkGeo* cg, SpeciesDev& sp, int* err, void* scan_tmp,
                cudaStream_t st, bool fp32)
{
  (void)cg;
  const size_t n     = (size_t)g.nchunk * g.ncell * LANES;
  const int    nblk  = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  int32_t*     bsum  = reinterpret_cast<int32_t*>(scan_tmp);
  int32_t*     total = bsum + nblk;
  k_scan_reduce<<<nblk, SCAN_THREADS, 0, st>>>(sp.hist, n, bsum);
  NIX_LAUNCHED();
  k_scan_bsum<<<1, 1024, 0, st>>>(bsum, nblk, total);
  NIX_LAUNCHED();
  k_scan_apply<<<nblk, SCAN_THREADS, 0, st>>>(sp.hist, n, bsum, sp.start);
  NIX_LAUNCHED();
  k_chunk_bases<<<(g.nchunk + 1 + 255) / 256, 256, 0, st>>>(g, sp, total, err);
  NIX_LAUNCHED();
  k_place<<<grid_for(sp.cap, 256), 256, 0, st>>>(g, sp, err);
  NIX_LAUNCHED();
  k_place_msg<<<grid_for(sp.lcap, 256, 148 * 4), 256, 0, st>>>(g, sp, err);
  NIX_LAUNCHED();
  if (fp32) k_gather<float><<<grid_for(sp.cap, 256), 256, 0, st>>>(g, sp, err);
  else k_gather<double><<<grid_for(sp.cap, 256), 256, 0, st>>>(g, sp, err);
  NIX_LAUNCHED();
  return 0;
}
} // namespace nixb200
