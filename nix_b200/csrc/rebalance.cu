// rebalance.cu -- chunks in the reference's wire format, produced and consumed ON THE DEVICE, and the
// device-to-device rebalance built on it (sm_100a): SURVEY.md section 8f, row N2
//
// The reference rebalances by Balancer::sendrecv_chunk (balancer.hpp:122-332): every chunk whose owner
// changed is serialised with Chunk::pack (chunk.cpp:18-60) -- which for a PIC chunk appends its field arrays
// and, per species, XtensorParticle::pack (xtensor_particle.hpp:128-169) -- shipped to rank-1 or rank+1 over
// MPI and rebuilt by the factory + unpack.  Behind the C ABI that used to mean: download every chunk of the
// rank, pack on the host, MPI, rebuild and re-upload the whole domain.  Here:
//
//   nixb200_chunk_wire_size / _pack   the bytes that follow nix::Chunk::pack's own header in such a record --
//        order, ns, uf, uj, then per species XtensorParticle::pack: 29 scalars (175 bytes, unaligned!), xu
//        [Np_total][7], xv, gindex, pindex [Ng+1], pcount [Ng+1][8] -- assembled on the device (transposed SoA
//        -> AoS, pindex / pcount expanded to the reference's layout) into device or host memory.  A record made
//        of the reference's header + this payload is read by the reference's own unpack (checked in
//        host/demo_main.cpp against XtensorParticle::unpack).
//   nixb200_domain_rebalance          collective over the ranks: given the new rank boundaries (from the
//        host's unchanged Balancer::assign), the chunks that change owner travel in that format from GPU to
//        GPU over NCCL -- to rank-1 / rank+1 only, like the reference -- while the chunks that stay are moved
//        device-to-device into the re-sized arrays.  No particle or field byte crosses PCIe.
#include "domain.hpp"

#include <algorithm>
#include <cstring>

namespace nixb200
{
namespace
{
constexpr size_t PHDR = 3 * 4 + 2 * 8 + 3 + 6 * 4 + 15 * 8; // scalars of XtensorParticle::pack: 175 bytes

struct WireSpecies {
  size_t hdr, xu, xv, gindex, pindex, pcount;
  int    np, np_total;
};
struct WireLayout {
  size_t                   total, order, ns, uf, uj;
  std::vector<WireSpecies> sp;
};

WireLayout wire_layout(size_t cells, size_t ns, const int32_t* np);
inline WireLayout wire_layout(const Domain* d, const int32_t* np)
{
  return wire_layout(d->cells_per_chunk, d->sp.size(), np);
}

int round_up_alloc(int np)
{
  return ((np + 128) / 128) * 128; // particle.hpp:146-153
}

WireLayout wire_layout(size_t cells, size_t ns, const int32_t* np)
{
  WireLayout w;
  size_t     a = 0;
  w.order = a, a += 4;
  w.ns = a, a += 4;
  w.uf = a, a += cells * 6 * 8;
  w.uj = a, a += cells * 4 * 8;
  for (size_t is = 0; is < ns; is++) {
    WireSpecies s;
    s.np       = np[is];
    s.np_total = round_up_alloc(np[is]);
    s.hdr = a, a += PHDR;
    s.xu = a, a += (size_t)s.np_total * NC * 8;
    s.xv = a, a += (size_t)s.np_total * NC * 8;
    s.gindex = a, a += (size_t)s.np_total * 4;
    s.pindex = a, a += (cells + 1) * 4;
    s.pcount = a, a += (cells + 1) * LANES * 4;
    w.sp.push_back(s);
  }
  w.total = a;
  return w;
}

// the 175 scalar bytes of XtensorParticle::pack (xtensor_particle.hpp:130-159) from plain numbers (host logic):
// dims / del / origin / glo / ghi in (z, y, x) order; origin = the chunk's lower corner, offset * del (chunk.cpp:217-232)
void particle_header_plain(const int* dims, int nb, const double* del, const double* origin, const double* glo,
                           const double* ghi, double q, double m, int np, int np_total, unsigned char* out)
{
  unsigned char* p = out;
  auto put = [&](const void* v, size_t n) {
    std::memcpy(p, v, n);
    p += n;
  };
  const int  Ng  = (dims[0] + 2 * nb) * (dims[1] + 2 * nb) * (dims[2] + 2 * nb);
  const bool yes = true;
  put(&np_total, 4), put(&np, 4), put(&Ng, 4), put(&q, 8), put(&m, 8);
  put(&yes, 1), put(&yes, 1), put(&yes, 1);
  for (int a = 2; a >= 0; a--) { // Lbx Ubx Lby Uby Lbz Ubz
    const int lb = nb, ub = nb + dims[a] - 1;
    put(&lb, 4), put(&ub, 4);
  }
  for (int a = 2; a >= 0; a--) put(&del[a], 8); // delx dely delz
  for (int a = 2; a >= 0; a--) {                // xmin xmax ymin ymax zmin zmax (chunk.cpp:217-232)
    const double lo = origin[a], hi = origin[a] + dims[a] * del[a];
    put(&lo, 8), put(&hi, 8);
  }
  for (int a = 2; a >= 0; a--) put(&glo[a], 8), put(&ghi[a], 8);
}

// ... for species `is` of local chunk k
void particle_header(const Domain* d, int k, int is, const WireSpecies& ws, unsigned char* out)
{
  const Geo& g = d->geo;
  particle_header_plain(g.N, g.nb, g.del, &d->origin_host[3 * k], g.glo, g.ghi, d->sp[is].q, d->sp[is].m, ws.np, ws.np_total,
                        out);
}

// pindex [Ng+1] / pcount [Ng+1][8] of one chunk in the reference's layout from the device's compact scan
__global__ void k_wire_index(const int32_t* __restrict__ start, const int32_t* __restrict__ oob8, int ncell, int Ng,
                             int32_t* __restrict__ pindex, int32_t* __restrict__ pcount)
{
  const int base = start[0], npk = start[(size_t)ncell * LANES] - base;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < (Ng + 1) * LANES; t += gridDim.x * blockDim.x) {
    const int ii = t / LANES, l = t % LANES;
    if (l == 0) pindex[ii] = (ii < ncell) ? start[(size_t)ii * LANES] - base : npk;
    pcount[t] = (ii < ncell) ? start[t + 1] - start[t] : ((ii == Ng) ? oob8[l] : 0);
  }
}

__global__ void k_zero_bytes(unsigned char* p, size_t n)
{
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) p[t] = 0;
}

// byte copy between arbitrarily aligned device addresses (cudaMemcpyAsync would do; a kernel keeps hundreds
// of small pieces of a rebalance in stream order without host round trips)
int dev_copy(void* dst, const void* src, size_t n, cudaStream_t st)
{
  if (n) NIX_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, st));
  return 0;
}

// assemble the payload of local chunk k at `wire` (DEVICE memory, w.total bytes).  cb[is] = first particle of
// the chunk in the species store.  Uses d->pack_dev as aligned staging.
int wire_pack_chunk(Domain* d, int k, const WireLayout& w, const int32_t* cb, unsigned char* wire)
{
  const Geo&   g     = d->geo;
  const size_t cells = d->cells_per_chunk;
  const int    order = g.order, ns = (int)d->sp.size();
  size_t       stage = std::max(cells * 6 * 8, (cells + 1) * (LANES + 1) * 4);
  for (auto& s : w.sp) stage = std::max(stage, (size_t)s.np_total * NC * 8);
  if (pack_scratch(d, stage)) return 1;
  unsigned char hdr[8];
  std::memcpy(hdr, &order, 4);
  std::memcpy(hdr + 4, &ns, 4);
  NIX_CUDA(cudaMemcpyAsync(wire + w.order, hdr, 8, cudaMemcpyHostToDevice, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream)); // (hdr is a stack buffer)
  for (int which = 0; which < 2; which++) {
    const int    nc  = which == 0 ? 6 : 4, fc = which == 0 ? d->fcs : 4;
    const char*  src = reinterpret_cast<const char*>(which == 0 ? d->uf : d->uj) + (size_t)k * cells * fc * d->esz;
    const size_t off = which == 0 ? w.uf : w.uj;
    if (d->fp32) {
      if (launch_cells_convert(false, d->pack_dev, reinterpret_cast<float*>(const_cast<char*>(src)), cells, nc, fc, d->stream)) return 1;
      if (dev_copy(wire + off, d->pack_dev, cells * nc * 8, d->stream)) return 1;
    } else if (dev_copy(wire + off, src, cells * nc * 8, d->stream)) return 1;
  }
  for (int is = 0; is < ns; is++) {
    const WireSpecies& s  = w.sp[is];
    SpeciesDev&        sd = d->sp[is];
    unsigned char      ph[PHDR];
    particle_header(d, k, is, s, ph);
    NIX_CUDA(cudaMemcpyAsync(wire + s.hdr, ph, PHDR, cudaMemcpyHostToDevice, d->stream));
    NIX_CUDA(cudaStreamSynchronize(d->stream));
    // xu: [Np_total][7], the chunk's particles in container (cell-sorted) order, zero beyond Np
    k_zero_bytes<<<148, 256, 0, d->stream>>>(reinterpret_cast<unsigned char*>(d->pack_dev), (size_t)s.np_total * NC * 8);
    NIX_LAUNCHED();
    if (d->fp32) {
      if (launch_soa_to_aos_f32(reinterpret_cast<const float*>(sd.xu), d->pack_dev, sd.cap, (size_t)cb[is], (size_t)s.np,
                                d->origin_dev + 3 * k, d->stream))
        return 1;
    } else if (launch_soa_to_aos(sd.xu, d->pack_dev, sd.cap, (size_t)cb[is], (size_t)s.np, d->stream)) return 1;
    if (dev_copy(wire + s.xu, d->pack_dev, (size_t)s.np_total * NC * 8, d->stream)) return 1;
    // xv (the temporary array) and gindex are scratch in the reference as well: zeros
    k_zero_bytes<<<148, 256, 0, d->stream>>>(wire + s.xv, (size_t)s.np_total * NC * 8 + (size_t)s.np_total * 4);
    NIX_LAUNCHED();
    int32_t* pidx = reinterpret_cast<int32_t*>(d->pack_dev);
    int32_t* pcnt = pidx + (cells + 1);
    k_wire_index<<<64, 256, 0, d->stream>>>(sd.start + (size_t)k * g.ncell * LANES, sd.oob + (size_t)k * LANES, g.ncell,
                                            (int)cells, pidx, pcnt);
    NIX_LAUNCHED();
    if (dev_copy(wire + s.pindex, pidx, (cells + 1) * 4, d->stream)) return 1;
    if (dev_copy(wire + s.pcount, pcnt, (cells + 1) * LANES * 4, d->stream)) return 1;
  }
  return 0;
}

// consume a payload: fields into chunk k of `nd`, particles of species is at position first[is] of its store
int wire_unpack_chunk(Domain* nd, int k, const WireLayout& w, const unsigned char* wire, const int32_t* first)
{
  const size_t cells = nd->cells_per_chunk;
  size_t       stage = cells * 6 * 8;
  for (auto& s : w.sp) stage = std::max(stage, (size_t)std::max(1, s.np) * NC * 8);
  if (pack_scratch(nd, stage)) return 1;
  for (int which = 0; which < 2; which++) {
    const int    nc  = which == 0 ? 6 : 4, fc = which == 0 ? nd->fcs : 4;
    char*        dst = reinterpret_cast<char*>(which == 0 ? nd->uf : nd->uj) + (size_t)k * cells * fc * nd->esz;
    const size_t off = which == 0 ? w.uf : w.uj;
    if (nd->fp32) {
      if (dev_copy(nd->pack_dev, wire + off, cells * nc * 8, nd->stream)) return 1;
      if (launch_cells_convert(true, nd->pack_dev, reinterpret_cast<float*>(dst), cells, nc, fc, nd->stream)) return 1;
    } else if (dev_copy(dst, wire + off, cells * nc * 8, nd->stream)) return 1;
  }
  for (size_t is = 0; is < nd->sp.size(); is++) {
    const WireSpecies& s  = w.sp[is];
    SpeciesDev&        sd = nd->sp[is];
    // per-(cell, lane) counts as the sender's last count() left them (pcount rows [0, ncell) and row Ng)
    const size_t nk = (size_t)nd->geo.ncell * LANES;
    if (dev_copy(sd.hist + (size_t)k * nk, wire + s.pcount, nk * 4, nd->stream)) return 1;
    if (dev_copy(sd.oob + (size_t)k * LANES, wire + s.pcount + cells * LANES * 4, LANES * 4, nd->stream)) return 1;
    if (s.np == 0) continue;
    if (dev_copy(nd->pack_dev, wire + s.xu, (size_t)s.np * NC * 8, nd->stream)) return 1;
    if (nd->fp32) {
      const Geo&   g = nd->geo;
      const double extent[3] = {g.N[0] * g.del[0], g.N[1] * g.del[1], g.N[2] * g.del[2]};
      if (launch_aos_to_soa_f32(nd->pack_dev, reinterpret_cast<float*>(sd.xu), sd.cap, (size_t)first[is], (size_t)s.np, sd.cbase,
                                g.nchunk, nd->origin_dev, extent, nd->stream))
        return 1;
    } else if (launch_aos_to_soa(nd->pack_dev, sd.xu, sd.cap, (size_t)first[is], (size_t)s.np, nd->stream)) return 1;
  }
  return 0;
}

int read_cbase(Domain* d, std::vector<std::vector<int32_t>>& cb)
{
  const int nch = d->geo.nchunk;
  cb.assign(d->sp.size(), std::vector<int32_t>(nch + 1));
  for (size_t is = 0; is < d->sp.size(); is++)
    NIX_CUDA(cudaMemcpyAsync(cb[is].data(), d->sp[is].cbase, sizeof(int32_t) * (nch + 1), cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  return 0;
}
} // namespace
} // namespace nixb200

using namespace nixb200;

extern "C" {

int nixb200_chunk_wire_size(nixb200_domain* dd, int k, int64_t* bytes)
{
  Domain* d = reinterpret_cast<Domain*>(dd);
  if (!d || !bytes || k < 0 || k >= d->geo.nchunk) {
    set_error("chunk_wire_size: bad argument");
    return 1;
  }
  DeviceGuard                       dev_guard(d->desc.device);
  std::vector<std::vector<int32_t>> cb;
  if (read_cbase(d, cb)) return 1;
  std::vector<int32_t> np(d->sp.size());
  for (size_t is = 0; is < np.size(); is++) np[is] = cb[is][k + 1] - cb[is][k];
  *bytes = (int64_t)wire_layout(d, np.data()).total;
  return 0;
}

// the same from the chunk shape and the particle counts alone: host logic, no device
int nixb200_wire_size_dims(const int* dims, int nb, int ns, const int* np, int64_t* bytes)
{
  if (!dims || !np || !bytes || nb < 0 || ns < 1 || dims[0] < 1 || dims[1] < 1 || dims[2] < 1) {
    set_error("wire_size_dims: bad argument");
    return 1;
  }
  for (int is = 0; is < ns; is++)
    if (np[is] < 0) {
      set_error("wire_size_dims: negative particle count");
      return 1;
    }
  const size_t cells = (size_t)(dims[0] + 2 * nb) * (dims[1] + 2 * nb) * (dims[2] + 2 * nb);
  *bytes             = (int64_t)wire_layout(cells, (size_t)ns, np).total;
  return 0;
}

// the 175 scalar bytes that open a species' part of the record (host logic): offset / gdims in cells as
// Chunk::set_global_context takes them (chunk.cpp:239-247), all triples in (z, y, x) order
int nixb200_wire_particle_header(const int* dims, int nb, const double* del, const int* offset, const int* gdims, double q,
                                 double m, int np, void* out175)
{
  if (!dims || !del || !offset || !gdims || !out175 || nb < 0 || np < 0) {
    set_error("wire_particle_header: bad argument");
    return 1;
  }
  double origin[3], glo[3], ghi[3];
  for (int a = 0; a < 3; a++) {
    origin[a] = offset[a] * del[a]; // domain.cu: origin_host, cg.lo
    glo[a]    = 0.0;
    ghi[a]    = gdims[a] * del[a]; // domain.cu: g.ghi = (cdims * dims) * del
  }
  particle_header_plain(dims, nb, del, origin, glo, ghi, q, m, np, round_up_alloc(np), reinterpret_cast<unsigned char*>(out175));
  return 0;
}

int nixb200_chunk_wire_pack(nixb200_domain* dd, int k, void* buffer, int64_t bytes)
{
  Domain* d = reinterpret_cast<Domain*>(dd);
  if (!d || !buffer || k < 0 || k >= d->geo.nchunk) {
    set_error("chunk_wire_pack: bad argument");
    return 1;
  }
  DeviceGuard                       dev_guard(d->desc.device);
  std::vector<std::vector<int32_t>> cb;
  if (read_cbase(d, cb)) return 1;
  std::vector<int32_t> np(d->sp.size()), first(d->sp.size());
  for (size_t is = 0; is < np.size(); is++) np[is] = cb[is][k + 1] - cb[is][k], first[is] = cb[is][k];
  WireLayout w = wire_layout(d, np.data());
  if ((int64_t)w.total > bytes) {
    set_error("chunk_wire_pack: buffer too small");
    return 1;
  }
  unsigned char* wire = nullptr;
  NIX_CUDA(cudaMalloc(&wire, w.total));
  int rc = wire_pack_chunk(d, k, w, first.data(), wire);
  if (!rc && cudaMemcpyAsync(buffer, wire, w.total, cudaMemcpyDefault, d->stream) != cudaSuccess) rc = 1;
  cudaStreamSynchronize(d->stream);
  cudaFree(wire);
  return rc;
}

// Pure host logic (no device): which chunk ids a rank that owned [b0, e0) and will own [b1, e1) sends to rank-1
// ([sl0, sl1)) and to rank+1 ([sr0, sr1)), receives from rank-1 ([rl0, rl1)) and from rank+1 ([rr0, rr1)), and
// keeps ([keep0, keep1)); out[10] in that order.  Returns non-zero when the two ranges do not overlap.
int nixb200_rebalance_moves(int b0, int e0, int b1, int e1, int* out)
{
  if (!out) {
    set_error("rebalance_moves: null output");
    return 1;
  }
  out[0] = b0, out[1] = std::max(b0, std::min(b1, e0));  // leaving at the low end
  out[2] = std::min(e0, std::max(e1, b0)), out[3] = e0;  // leaving at the high end
  out[4] = b1, out[5] = std::max(b1, std::min(b0, e1));  // arriving at the low end (ids below b0)
  out[6] = std::min(e1, std::max(e0, b1)), out[7] = e1;  // arriving at the high end
  out[8] = std::max(b0, b1), out[9] = std::min(e0, e1);  // staying
  return out[8] < out[9] ? 0 : 1;
}

// Collective: every rank of the communicator calls it with the same new boundaries.
int nixb200_domain_rebalance(nixb200_domain* dd, int nrank, const int* boundary, int rank)
{
  Domain* d = reinterpret_cast<Domain*>(dd);
  if (!d || !boundary || nrank < 1 || rank < 0 || rank >= nrank) {
    set_error("domain_rebalance: bad argument");
    return 1;
  }
  DeviceGuard dev_guard(d->desc.device);
  const int   b0 = d->desc.id_begin, e0 = d->desc.id_end, b1 = boundary[rank], e1 = boundary[rank + 1];
  if (e1 <= b1) {
    set_error("domain_rebalance: a rank must keep at least one chunk");
    return 1;
  }
  // the reference ships between neighbouring ranks only (balancer.hpp:130-133): what leaves at the low end goes to
  // rank-1, what leaves at the high end to rank+1, and at least one chunk stays where it is
  if (std::max(b0, b1) >= std::min(e0, e1)) {
    set_error("domain_rebalance: the old and the new chunk range of a rank must overlap (chunks move to rank-1 / rank+1 only)");
    return 1;
  }
  if (nrank > 1 && !peer_comm(d)) {
    set_error("domain_rebalance: call nixb200_domain_set_ranks and comm_init / set_comm first");
    return 1;
  }
  const int ns = (int)d->sp.size();
  // moving chunks, in ascending id order: [send to left][send to right], [recv from left][recv from right]
  int mv[10];
  nixb200_rebalance_moves(b0, e0, b1, e1, mv);
  const int sl0 = mv[0], sl1 = mv[1], sr0 = mv[2], sr1 = mv[3], rl0 = mv[4], rl1 = mv[5], rr0 = mv[6], rr1 = mv[7];
  const int nsl = std::max(0, sl1 - sl0), nsr = std::max(0, sr1 - sr0), nrl = std::max(0, rl1 - rl0), nrr = std::max(0, rr1 - rr0);
  const int keep0 = mv[8], keep1 = mv[9];

  std::vector<std::vector<int32_t>> cb;
  if (read_cbase(d, cb)) return 1;
  auto np_of = [&](int id, int is) { return cb[is][id - b0 + 1] - cb[is][id - b0]; };

  // 1. particle counts of the moving chunks (the reference's size exchange, balancer.hpp:249-256)
  const int        left = rank > 0 ? rank - 1 : -1, right = rank < nrank - 1 ? rank + 1 : -1;
  std::vector<int32_t> cnt_s((size_t)(nsl + nsr) * ns + 1), cnt_r((size_t)(nrl + nrr) * ns + 1);
  for (int j = 0; j < nsl; j++)
    for (int is = 0; is < ns; is++) cnt_s[(size_t)j * ns + is] = np_of(sl0 + j, is);
  for (int j = 0; j < nsr; j++)
    for (int is = 0; is < ns; is++) cnt_s[(size_t)(nsl + j) * ns + is] = np_of(sr0 + j, is);
  if ((nsl && left < 0) || (nsr && right < 0) || (nrl && left < 0) || (nrr && right < 0)) {
    set_error("domain_rebalance: boundary[0] and boundary[nrank] are fixed");
    return 1;
  }
  int32_t *cs_dev = nullptr, *cr_dev = nullptr;
  NIX_CUDA(cudaMalloc(&cs_dev, sizeof(int32_t) * cnt_s.size()));
  NIX_CUDA(cudaMalloc(&cr_dev, sizeof(int32_t) * cnt_r.size()));
  NIX_CUDA(cudaMemcpyAsync(cs_dev, cnt_s.data(), sizeof(int32_t) * cnt_s.size(), cudaMemcpyHostToDevice, d->stream));
  if (nrank > 1) {
    const int    ranks[2]  = {left, right};
    const void*  sb[2]     = {cs_dev, cs_dev + (size_t)nsl * ns};
    const size_t sbytes[2] = {sizeof(int32_t) * nsl * ns, sizeof(int32_t) * nsr * ns};
    void*        rb[2]     = {cr_dev, cr_dev + (size_t)nrl * ns};
    const size_t rbytes[2] = {sizeof(int32_t) * nrl * ns, sizeof(int32_t) * nrr * ns};
    if (peer_sendrecv_bytes(peer_comm(d), d->stream, 2, ranks, sb, sbytes, rb, rbytes)) return 1;
  }
  NIX_CUDA(cudaMemcpyAsync(cnt_r.data(), cr_dev, sizeof(int32_t) * cnt_r.size(), cudaMemcpyDeviceToHost, d->stream));
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  cudaFree(cs_dev);
  cudaFree(cr_dev);

  // 2. wire records of the leaving chunks, one contiguous device buffer per neighbour
  std::vector<WireLayout> ws, wr;
  size_t                  tot_s[2] = {0, 0}, tot_r[2] = {0, 0};
  for (int j = 0; j < nsl + nsr; j++) {
    ws.push_back(wire_layout(d, &cnt_s[(size_t)j * ns]));
    tot_s[j < nsl ? 0 : 1] += ws.back().total;
  }
  for (int j = 0; j < nrl + nrr; j++) {
    wr.push_back(wire_layout(d, &cnt_r[(size_t)j * ns]));
    tot_r[j < nrl ? 0 : 1] += wr.back().total;
  }
  unsigned char *sbuf = nullptr, *rbuf = nullptr;
  NIX_CUDA(cudaMalloc(&sbuf, std::max<size_t>(1, tot_s[0] + tot_s[1])));
  NIX_CUDA(cudaMalloc(&rbuf, std::max<size_t>(1, tot_r[0] + tot_r[1])));
  {
    size_t off = 0;
    for (int j = 0; j < nsl + nsr; j++) {
      const int            id = j < nsl ? sl0 + j : sr0 + (j - nsl);
      std::vector<int32_t> first(ns);
      for (int is = 0; is < ns; is++) first[is] = cb[is][id - b0];
      if (wire_pack_chunk(d, id - b0, ws[j], first.data(), sbuf + off)) return 1;
      off += ws[j].total;
    }
  }
  if (nrank > 1) {
    const int    ranks[2]  = {left, right};
    const void*  sb[2]     = {sbuf, sbuf + tot_s[0]};
    void*        rb[2]     = {rbuf, rbuf + tot_r[0]};
    if (peer_sendrecv_bytes(peer_comm(d), d->stream, 2, ranks, sb, tot_s, rb, tot_r)) return 1;
  }

  // 3. the re-sized domain: same streams and communicator, new id range.  Peak memory = the old and the new
  //    particle arrays (xu) + the wire buffers: everything that carries no state between steps (temporary array,
  //    keys, member lists, migration buffers) is freed first and re-created last
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  for (auto& s : d->sp) free_species_scratch(s);
  nixb200_domain_desc nd_desc = d->desc;
  nd_desc.id_begin            = b1;
  nd_desc.id_end              = e1;
  std::vector<double> q(ns), m(ns);
  for (int is = 0; is < ns; is++) q[is] = d->sp[is].q, m[is] = d->sp[is].m;
  nixb200_domain* ndh = nullptr;
  if (nixb200_domain_create(&nd_desc, d->coord_all.data(), q.data(), m.data(), &ndh)) return 1;
  Domain* nd = reinterpret_cast<Domain*>(ndh);
  // run everything below on the OLD domain's stream so that the two domains are ordered
  cudaStream_t nd_own = nd->stream;
  nd->stream          = d->stream;

  const int nch1 = e1 - b1;
  for (int is = 0; is < ns; is++) {
    // new chunk bases: [arrivals low][kept][arrivals high]
    std::vector<int32_t> nb(nch1 + 1, 0);
    for (int id = b1; id < e1; id++) {
      int n;
      if (id < keep0) n = cnt_r[(size_t)(id - rl0) * ns + is];
      else if (id < keep1) n = np_of(id, is);
      else n = cnt_r[(size_t)(nrl + (id - rr0)) * ns + is];
      nb[id - b1 + 1] = nb[id - b1] + n;
    }
    SpeciesDev& s = nd->sp[is];
    if (alloc_species_particles(nd, s, nb[nch1], /*with_scratch=*/false)) return 1;
    NIX_CUDA(cudaMemcpyAsync(s.cbase, nb.data(), sizeof(int32_t) * (nch1 + 1), cudaMemcpyHostToDevice, d->stream));
    NIX_CUDA(cudaStreamSynchronize(d->stream)); // (nb is a local)
    // kept chunks: one contiguous particle range per component, device to device
    const int32_t src0 = cb[is][keep0 - b0], nkeep = cb[is][keep1 - b0] - src0, dst0 = nb[keep0 - b1];
    for (int c = 0; c < d->nct; c++)
      NIX_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(s.xu) + ((size_t)c * s.cap + dst0) * d->esz,
                               reinterpret_cast<char*>(d->sp[is].xu) + ((size_t)c * d->sp[is].cap + src0) * d->esz,
                               (size_t)nkeep * d->esz, cudaMemcpyDeviceToDevice, d->stream));
    // ... and their per-(cell, lane) counts and out-of-bounds rows (NOT a re-sort: the key of the reference's sort
    // is cell * 8 + position % 8, so sorting a sorted container regroups the lanes)
    const size_t nk = (size_t)d->geo.ncell * LANES;
    if (launch_hist_from_start(d->sp[is].start + (size_t)(keep0 - b0) * nk, s.hist + (size_t)(keep0 - b1) * nk,
                               (size_t)(keep1 - keep0) * nk, d->stream))
      return 1;
    NIX_CUDA(cudaMemcpyAsync(s.oob + (size_t)(keep0 - b1) * LANES, d->sp[is].oob + (size_t)(keep0 - b0) * LANES,
                             sizeof(int32_t) * LANES * (keep1 - keep0), cudaMemcpyDeviceToDevice, d->stream));
  }
  for (int which = 0; which < 2; which++) { // kept chunks: E/B and J
    const size_t cb_ = d->cells_per_chunk * (which == 0 ? d->fcs : 4) * d->esz;
    NIX_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(which == 0 ? nd->uf : nd->uj) + (size_t)(keep0 - b1) * cb_,
                             reinterpret_cast<char*>(which == 0 ? d->uf : d->uj) + (size_t)(keep0 - b0) * cb_,
                             (size_t)(keep1 - keep0) * cb_, cudaMemcpyDeviceToDevice, d->stream));
  }
  {
    // arrivals
    std::vector<std::vector<int32_t>> ncb;
    if (read_cbase(nd, ncb)) return 1;
    size_t off = 0;
    for (int j = 0; j < nrl + nrr; j++) {
      const int            id = j < nrl ? rl0 + j : rr0 + (j - nrl);
      std::vector<int32_t> first(ns);
      for (int is = 0; is < ns; is++) first[is] = ncb[is][id - b1];
      if (wire_unpack_chunk(nd, id - b1, wr[j], rbuf + off, first.data())) return 1;
      off += wr[j].total;
    }
  }
  // scan of the counts -> start / chunk bases of the new domain (the particles are where they belong already)
  for (int is = 0; is < ns; is++) {
    SpeciesDev& s = nd->sp[is];
    if (launch_scan_only(nd->geo, s, nd->err_dev, nd->scan_tmp, d->stream)) return 1;
    std::swap(s.cbase, s.cbase_new); // (equal to the bases laid out above)
  }
  NIX_CUDA(cudaStreamSynchronize(d->stream));
  cudaFree(sbuf);
  cudaFree(rbuf);

  // 4. the caller's handle takes over the new arrays; streams, profiling state and the communicator stay
  void*      comm = nullptr;
  const bool own  = peer_take_comm(d, &comm);
  nd->stream      = nd_own;
  std::swap(*d, *nd);
  std::swap(d->stream, nd->stream);
  std::swap(d->owns_stream, nd->owns_stream);
  std::swap(d->copy_stream, nd->copy_stream);
  std::swap(d->profiling, nd->profiling);
  for (int w = 0; w < 2; w++) {
    std::swap(d->ev_main_done[w], nd->ev_main_done[w]);
    std::swap(d->ev_copy_done[w], nd->ev_copy_done[w]);
  }
  nixb200_domain_destroy(reinterpret_cast<nixb200_domain*>(nd)); // frees the old arrays
  for (auto& s : d->sp)
    if (alloc_species_scratch(d, s)) return 1;
  d->particles_set = true;
  if (nixb200_domain_set_ranks(dd, nrank, boundary, rank)) return 1;
  if (comm && peer_give_comm(d, comm, own)) return 1;
  return 0;
}

} // extern "C"
