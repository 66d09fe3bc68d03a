// halo.cu -- E/B and J ghost-cell exchange between chunks (sm_100a)
//
// Replaces Chunk::{pack,begin,end,unpack}_bc_exchange (chunk.hpp:435-586) with
//   XtensorHaloField3D    copy   interior slab of the neighbour -> my ghost slab   xtensor_halo3d.hpp:28-70
//   XtensorHaloCurrent3D  add    ghost slab of the neighbour    -> my interior slab xtensor_halo3d.hpp:86-128
// for every pair of chunks living on the same device, without any intermediate buffer: each thread
// owns one destination cell and GATHERS from the neighbour chunk.  For the current the gather runs
// over the 26 directions in the reference's unpack order (iz, iy, ix ascending, chunk.hpp:480-499),
// so every interior cell sees the same sequence of additions as the reference: bit-exact, no atomics.
//
// Slab tables (chunk.cpp:171-207), per axis with Lb = nb, Ub = nb+N-1:
//   send[0] = [Lb, Lb+nb-1]   send[1] = [Lb, Ub]   send[2] = [Ub-nb+1, Ub]
//   recv[0] = [Lb-nb, Lb-1]   recv[1] = [Lb, Ub]   recv[2] = [Ub+1, Ub+nb]
// A message sent in direction d is received in slot 26-d of the neighbour (chunk.hpp:532-554).
//
// The *_buf kernels move the same slabs to / from one contiguous buffer in the reference's
// MpiBuffer layout (chunk.cpp:257-286); they serve neighbours on other ranks and the drop-in
// per-chunk API.
#include "common.cuh"

#include <algorithm>

namespace nixb200
{
namespace
{
struct SlotTable {
  int addr[27]; // offset in doubles of each slot inside the buffer
};

__device__ __forceinline__ size_t cell_off(const Geo& g, int ch, int iz, int iy, int ix)
{
  return (((size_t)ch * g.M[0] + iz) * g.M[1] + iy) * g.M[2] + ix;
}

// slab of direction index e (0/1/2) along axis a: first array index and width (chunk.cpp:171-207)
__device__ __forceinline__ void slab_bounds(const Geo& g, int a, int e, bool recv, int& lo, int& n)
{
  const int Lb = g.nb, Ub = g.nb + g.N[a] - 1;
  if (e == 1) {
    lo = Lb;
    n  = g.N[a];
  } else if (recv) {
    lo = (e == 0) ? Lb - g.nb : Ub + 1;
    n  = g.nb;
  } else {
    lo = (e == 0) ? Lb : Ub - g.nb + 1;
    n  = g.nb;
  }
}

// position of array cell i inside the (row-major z,y,x) slab of slot `slot`
__device__ __forceinline__ size_t slab_index(const Geo& g, int slot, bool recv, const int* i)
{
  int lo[3], n[3];
  slab_bounds(g, 0, slot / 9, recv, lo[0], n[0]);
  slab_bounds(g, 1, (slot / 3) % 3, recv, lo[1], n[1]);
  slab_bounds(g, 2, slot % 3, recv, lo[2], n[2]);
  return ((size_t)(i[0] - lo[0]) * n[1] + (i[1] - lo[1])) * n[2] + (i[2] - lo[2]);
}

// ---- exchange: same-device neighbours are read in place, neighbours on other ranks from the
//      peer-major receive buffer (peer.cu) ----------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_halo_field(Geo g, const ChunkGeo* __restrict__ cg, T* uf, PeerTabs pt,
                                                    const T* __restrict__ recvbuf)
{
  const int ch    = blockIdx.y;
  const int ncell = g.M[0] * g.M[1] * g.M[2];
  const int Lb = g.nb;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ncell; t += gridDim.x * blockDim.x) {
    int ix = t % g.M[2];
    int iy = (t / g.M[2]) % g.M[1];
    int iz = t / (g.M[2] * g.M[1]);
    int i[3] = {iz, iy, ix};
    int e[3], s[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      int Ub = Lb + g.N[a] - 1;
      e[a]   = (i[a] < Lb) ? 0 : ((i[a] > Ub) ? 2 : 1);
      s[a]   = i[a] + (1 - e[a]) * g.N[a]; // ghost low <- neighbour's high interior, and vice versa
    }
    int slot = 9 * e[0] + 3 * e[1] + e[2];
    if (slot == 13) continue;
    int nb = cg[ch].nbr[slot];
    // one cell = 48 bytes (6 doubles) or 32 bytes (8 floats): whole 16-byte words either way
    constexpr int FC = field_stride<T>(), NV = FC * (int)sizeof(T) / 16;
    const int4* src;
    if (nb >= 0) {
      src = reinterpret_cast<const int4*>(uf + cell_off(g, nb, s[0], s[1], s[2]) * FC);
    } else {
      int j = (pt.recv_slot != nullptr) ? pt.recv_slot[ch * 27 + slot] : -1;
      if (j < 0) continue;
      src = reinterpret_cast<const int4*>(recvbuf + ((size_t)pt.recv_ent[j].celloff + slab_index(g, slot, true, i)) * FC);
    }
    int4* dst = reinterpret_cast<int4*>(uf + cell_off(g, ch, iz, iy, ix) * FC);
    int4  v[NV];
#pragma unroll
    for (int q = 0; q < NV; q++) v[q] = src[q];
#pragma unroll
    for (int q = 0; q < NV; q++) dst[q] = v[q];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_halo_current(Geo g, const ChunkGeo* __restrict__ cg, T* uj, PeerTabs pt,
                                                      const T* __restrict__ recvbuf)
{
  const int ch    = blockIdx.y;
  const int ncell = g.N[0] * g.N[1] * g.N[2];
  const int Lb = g.nb, nbw = g.nb;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ncell; t += gridDim.x * blockDim.x) {
    int i[3];
    i[2] = t % g.N[2] + Lb;
    i[1] = (t / g.N[2]) % g.N[1] + Lb;
    i[0] = t / (g.N[2] * g.N[1]) + Lb;
    bool lowm[3], higm[3];
    bool any = false;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      int Ub  = Lb + g.N[a] - 1;
      lowm[a] = i[a] <= Lb + nbw - 1;
      higm[a] = i[a] >= Ub - nbw + 1;
      any     = any || lowm[a] || higm[a];
    }
    if (!any) continue;
    T* dst = uj + cell_off(g, ch, i[0], i[1], i[2]) * 4;
    T  v[4] = {dst[0], dst[1], dst[2], dst[3]};
    for (int slot = 0; slot < 27; slot++) {
      if (slot == 13) continue;
      int  e[3] = {slot / 9, (slot / 3) % 3, slot % 3};
      bool in   = true;
      int  s[3];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        in   = in && (e[a] == 1 || (e[a] == 0 && lowm[a]) || (e[a] == 2 && higm[a]));
        s[a] = i[a] + (1 - e[a]) * g.N[a]; // my low interior <- neighbour's high ghost, and vice versa
      }
      if (!in) continue;
      int nb = cg[ch].nbr[slot];
      const T* src;
      if (nb >= 0) {
        src = uj + cell_off(g, nb, s[0], s[1], s[2]) * 4;
      } else {
        int j = (pt.recv_slot != nullptr) ? pt.recv_slot[ch * 27 + slot] : -1;
        if (j < 0) continue;
        src = recvbuf + ((size_t)pt.recv_ent[j].celloff + slab_index(g, slot, false, i)) * 4;
      }
      // std::plus(buffer, cell)  xtensor_halo3d.hpp:125
#pragma unroll
      for (int c = 0; c < 4; c++) v[c] = add<true>(src[c], v[c]);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) dst[c] = v[c];
  }
}

// XtensorHaloMoment3D (xtensor_halo3d.hpp:135-187): the same gather-and-add as the current for an array with
// `ncomp` values per cell (ns * 14 moments); one thread per (interior cell, component)
template <typename T>
__global__ void __launch_bounds__(256) k_halo_moment(Geo g, const ChunkGeo* __restrict__ cg, T* um, int ncomp, PeerTabs pt,
                                                     const T* __restrict__ recvbuf)
{
  const int ch    = blockIdx.y;
  const int ncell = g.N[0] * g.N[1] * g.N[2];
  const int Lb = g.nb, nbw = g.nb;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)ncell * ncomp;
       t += (long long)gridDim.x * blockDim.x) {
    const int comp = (int)(t % ncomp), cidx = (int)(t / ncomp);
    int       i[3];
    i[2] = cidx % g.N[2] + Lb;
    i[1] = (cidx / g.N[2]) % g.N[1] + Lb;
    i[0] = cidx / (g.N[2] * g.N[1]) + Lb;
    bool lowm[3], higm[3], any = false;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      int Ub  = Lb + g.N[a] - 1;
      lowm[a] = i[a] <= Lb + nbw - 1;
      higm[a] = i[a] >= Ub - nbw + 1;
      any     = any || lowm[a] || higm[a];
    }
    if (!any) continue;
    T* dst = um + cell_off(g, ch, i[0], i[1], i[2]) * ncomp + comp;
    T  v   = *dst;
    for (int slot = 0; slot < 27; slot++) {
      if (slot == 13) continue;
      int  e[3] = {slot / 9, (slot / 3) % 3, slot % 3};
      bool in   = true;
      int  s[3];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        in   = in && (e[a] == 1 || (e[a] == 0 && lowm[a]) || (e[a] == 2 && higm[a]));
        s[a] = i[a] + (1 - e[a]) * g.N[a];
      }
      if (!in) continue;
      int nb = cg[ch].nbr[slot];
      if (nb >= 0) {
        v = add<true>(um[cell_off(g, nb, s[0], s[1], s[2]) * ncomp + comp], v);
      } else {
        int j = (pt.recv_slot != nullptr) ? pt.recv_slot[ch * 27 + slot] : -1;
        if (j < 0) continue;
        v = add<true>(recvbuf[((size_t)pt.recv_ent[j].celloff + slab_index(g, slot, false, i)) * ncomp + comp], v);
      }
    }
    *dst = v;
  }
}

// every slab bound for another rank -> the peer-major send buffer.  Field: SEND slabs (interior);
// current: RECV slabs (ghost, where the deposit spilled).  blockIdx.y = send entry.
template <typename T>
__global__ void __launch_bounds__(256) k_peer_pack(Geo g, int ncomp, bool recv_slab, const T* __restrict__ data,
                                                   const PeerEntry* __restrict__ ent, T* __restrict__ buf)
{
  const PeerEntry en = ent[blockIdx.y];
  int             lo[3], n[3];
  slab_bounds(g, 0, en.dir / 9, recv_slab, lo[0], n[0]);
  slab_bounds(g, 1, (en.dir / 3) % 3, recv_slab, lo[1], n[1]);
  slab_bounds(g, 2, en.dir % 3, recv_slab, lo[2], n[2]);
  const int total = en.cells * ncomp;
  T*        dst   = buf + (size_t)en.celloff * ncomp;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    int c  = t % ncomp;
    int r  = t / ncomp;
    int ix = r % n[2] + lo[2];
    int iy = (r / n[2]) % n[1] + lo[1];
    int iz = r / (n[2] * n[1]) + lo[0];
    dst[t] = data[cell_off(g, en.k, iz, iy, ix) * ncomp + c];
  }
}

// ---- buffer (MpiBuffer layout, chunk.cpp:257-286) pack / unpack of ONE chunk: the drop-in per-chunk API
// pack: field packs the SEND slabs (interior), current packs the RECV slabs (ghost)
__global__ void __launch_bounds__(256) k_halo_pack(Geo g, int ch, int ncomp, bool recv_slab,
                                                   const double* __restrict__ data, SlotTable tab,
                                                   double* __restrict__ buf)
{
  const int slot = blockIdx.y;
  if (slot == 13) return;
  int lo[3], n[3];
#pragma unroll
  for (int a = 0; a < 3; a++) slab_bounds(g, a, (a == 0) ? slot / 9 : (a == 1 ? (slot / 3) % 3 : slot % 3), recv_slab, lo[a], n[a]);
  const int total = n[0] * n[1] * n[2] * ncomp;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    int c  = t % ncomp;
    int r  = t / ncomp;
    int ix = r % n[2] + lo[2];
    int iy = (r / n[2]) % n[1] + lo[1];
    int iz = r / (n[2] * n[1]) + lo[0];
    buf[tab.addr[slot] + t] = data[cell_off(g, ch, iz, iy, ix) * ncomp + c];
  }
}

// unpack field: one thread per ghost value, reads the unique slot that covers it
__global__ void __launch_bounds__(256) k_halo_unpack_field(Geo g, int ch, double* __restrict__ uf,
                                                           SlotTable tab, const double* __restrict__ buf,
                                                           const int* __restrict__ nbvalid)
{
  const int ncell = g.M[0] * g.M[1] * g.M[2];
  const int Lb = g.nb;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ncell; t += gridDim.x * blockDim.x) {
    int i[3] = {t / (g.M[2] * g.M[1]), (t / g.M[2]) % g.M[1], t % g.M[2]};
    int e[3], lo[3], n[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      int Ub = Lb + g.N[a] - 1;
      e[a]   = (i[a] < Lb) ? 0 : ((i[a] > Ub) ? 2 : 1);
      slab_bounds(g, a, e[a], true, lo[a], n[a]);
    }
    int slot = 9 * e[0] + 3 * e[1] + e[2];
    if (slot == 13 || !nbvalid[slot]) continue;
    size_t r   = ((size_t)(i[0] - lo[0]) * n[1] + (i[1] - lo[1])) * n[2] + (i[2] - lo[2]);
    double* dst = uf + cell_off(g, ch, i[0], i[1], i[2]) * 6;
#pragma unroll
    for (int c = 0; c < 6; c++) dst[c] = buf[tab.addr[slot] + r * 6 + c];
  }
}

// unpack current: one thread per interior cell, adds the covering slots in the reference's order
__global__ void __launch_bounds__(256) k_halo_unpack_current(Geo g, int ch, double* __restrict__ uj,
                                                             SlotTable tab, const double* __restrict__ buf,
                                                             const int* __restrict__ nbvalid)
{
  const int ncell = g.N[0] * g.N[1] * g.N[2];
  const int Lb = g.nb;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ncell; t += gridDim.x * blockDim.x) {
    int i[3] = {t / (g.N[2] * g.N[1]) + Lb, (t / g.N[2]) % g.N[1] + Lb, t % g.N[2] + Lb};
    double* dst = uj + cell_off(g, ch, i[0], i[1], i[2]) * 4;
    double  v[4] = {dst[0], dst[1], dst[2], dst[3]};
    bool    touched = false;
    for (int slot = 0; slot < 27; slot++) {
      if (slot == 13 || !nbvalid[slot]) continue;
      int  e[3] = {slot / 9, (slot / 3) % 3, slot % 3};
      int  lo[3], n[3];
      bool in = true;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        slab_bounds(g, a, e[a], false, lo[a], n[a]);
        in = in && i[a] >= lo[a] && i[a] < lo[a] + n[a];
      }
      if (!in) continue;
      size_t r = ((size_t)(i[0] - lo[0]) * n[1] + (i[1] - lo[1])) * n[2] + (i[2] - lo[2]);
#pragma unroll
      for (int c = 0; c < 4; c++) v[c] = __dadd_rn(buf[tab.addr[slot] + r * 4 + c], v[c]);
      touched = true;
    }
    if (touched) {
#pragma unroll
      for (int c = 0; c < 4; c++) dst[c] = v[c];
    }
  }
}

SlotTable make_table(const Geo& g, int ncomp)
{
  SlotTable t;
  int       size = 0;
  for (int s = 0; s < 27; s++) {
    t.addr[s] = size;
    if (s == 13) continue;
    int e[3] = {s / 9, (s / 3) % 3, s % 3};
    int cnt  = ncomp;
    for (int a = 0; a < 3; a++) cnt *= (e[a] == 1) ? g.N[a] : g.nb;
    size += cnt;
  }
  return t;
}
} // namespace

int launch_halo_field(const Geo& g, const ChunkGeo* cg, void* uf, const PeerTabs& pt, const void* recvbuf,
                      cudaStream_t st, bool fp32)
{
  int  ncell = g.M[0] * g.M[1] * g.M[2];
  dim3 grid((ncell + 255) / 256, g.nchunk);
  if (fp32) k_halo_field<float><<<grid, 256, 0, st>>>(g, cg, (float*)uf, pt, (const float*)recvbuf);
  else k_halo_field<double><<<grid, 256, 0, st>>>(g, cg, (double*)uf, pt, (const double*)recvbuf);
  NIX_LAUNCHED();
  return 0;
}

int launch_halo_current(const Geo& g, const ChunkGeo* cg, void* uj, const PeerTabs& pt, const void* recvbuf,
                        cudaStream_t st, bool fp32)
{
  int  ncell = g.N[0] * g.N[1] * g.N[2];
  dim3 grid((ncell + 255) / 256, g.nchunk);
  if (fp32) k_halo_current<float><<<grid, 256, 0, st>>>(g, cg, (float*)uj, pt, (const float*)recvbuf);
  else k_halo_current<double><<<grid, 256, 0, st>>>(g, cg, (double*)uj, pt, (const double*)recvbuf);
  NIX_LAUNCHED();
  return 0;
}

int launch_halo_moment(const Geo& g, const ChunkGeo* cg, void* um, int ncomp, const PeerTabs& pt, const void* recvbuf,
                       cudaStream_t st, bool fp32)
{
  const long long n = (long long)g.N[0] * g.N[1] * g.N[2] * ncomp;
  dim3            grid((unsigned)std::min<long long>((n + 255) / 256, 4096), g.nchunk);
  if (fp32) k_halo_moment<float><<<grid, 256, 0, st>>>(g, cg, (float*)um, ncomp, pt, (const float*)recvbuf);
  else k_halo_moment<double><<<grid, 256, 0, st>>>(g, cg, (double*)um, ncomp, pt, (const double*)recvbuf);
  NIX_LAUNCHED();
  return 0;
}

// words per cell in the peer buffers: the device's own cell layout (E/B: 6 doubles or 8 floats; J: 4;
// moments: ncomp_moment = ns * 14)
int launch_peer_pack(const Geo& g, int mode, const void* data, const PeerTabs& pt, void* sendbuf,
                     cudaStream_t st, bool fp32, int ncomp_moment)
{
  if (pt.nsend == 0) return 0;
  const int ncomp = (mode == NIXB200_MODE_FIELD) ? (fp32 ? 8 : 6) : ((mode == NIXB200_MODE_MOMENT) ? ncomp_moment : 4);
  int       big   = g.nb * std::max(g.N[0], std::max(g.N[1], g.N[2])) * std::max(g.N[1], g.N[2]) * ncomp;
  dim3      grid(std::min(8, (big + 255) / 256), pt.nsend);
  if (fp32)
    k_peer_pack<float><<<grid, 256, 0, st>>>(g, ncomp, mode != NIXB200_MODE_FIELD, (const float*)data, pt.send_ent, (float*)sendbuf);
  else
    k_peer_pack<double><<<grid, 256, 0, st>>>(g, ncomp, mode != NIXB200_MODE_FIELD, (const double*)data, pt.send_ent, (double*)sendbuf);
  NIX_LAUNCHED();
  return 0;
}

int launch_halo_pack(const Geo& g, int k, int mode, const double* data, double* buf, cudaStream_t st)
{
  const int ncomp = (mode == NIXB200_MODE_FIELD) ? 6 : 4;
  SlotTable tab   = make_table(g, ncomp);
  int       big   = g.nb * g.N[1] * g.N[2] * ncomp;
  dim3      grid((big + 255) / 256, 27);
  k_halo_pack<<<grid, 256, 0, st>>>(g, k, ncomp, mode == NIXB200_MODE_CURRENT, data, tab, buf);
  NIX_LAUNCHED();
  return 0;
}

int launch_halo_unpack(const Geo& g, int k, int mode, double* data, const double* buf,
                       const int* nbvalid_dev, cudaStream_t st)
{
  if (mode == NIXB200_MODE_FIELD) {
    int ncell = g.M[0] * g.M[1] * g.M[2];
    k_halo_unpack_field<<<(ncell + 255) / 256, 256, 0, st>>>(g, k, data, make_table(g, 6), buf,
                                                             nbvalid_dev);
  } else {
    int ncell = g.N[0] * g.N[1] * g.N[2];
    k_halo_unpack_current<<<(ncell + 255) / 256, 256, 0, st>>>(g, k, data, make_table(g, 4), buf,
                                                               nbvalid_dev);
  }
  NIX_LAUNCHED();
  return 0;
}
} // namespace nixb200
