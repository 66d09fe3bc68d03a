// push_deposit.cu -- fused gather + Boris push + position update + Esirkepov deposit + count (sm_100a)
//
// Replaces, per particle (composition: oracle/ref/ref_driver.cpp):
//   digitize / shape_mc<O>      primitives.hpp:46-58,257-298,519-532
//   interp::shift_weights<O>    interp.hpp:149-172
//   interp::interp3d<O>         interp.hpp:95-113,217-230
//   push_boris, lorentz_factor  primitives.hpp:158-189
//   esirkepov::deposit3d<O>     esirkepov.hpp:155-237,326-340
//   append_current3d<O>         primitives.hpp:778-834
//   XtensorParticle::count      xtensor_particle.hpp:324-357   (of the NEW positions)
//   XtensorHaloParticle3D::pre_pack classification  xtensor_halo3d.hpp:288-302
//
// Work item = one (tz x ty x tx) tile of bins of one chunk, one CTA of 4 warps:
//   * TMA (cp.async.bulk.tensor.5d) stages the E/B tile of the CTA -- ghosts included -- into shared
//     memory; an mbarrier signals arrival.  A J tile of the CTA's deposit footprint lives in shared
//     memory next to it and is flushed to global memory ONCE per CTA with red.global.add.f64.
//   * a WARP owns one bin (cell) at a time; its particles are contiguous because the container is
//     cell-sorted, so the 32 lanes read 32 consecutive particles (coalesced SoA loads / stores).
//     Lane = particle: weights, gather from the E/B tile over the EXACT (O+1)^3 support of every
//     component, Boris, move, store, bin of the new position (-> key + histogram, or a leaver record
//     with its ordered rank inside the bin).
//   * deposit: every lane evaluates the Esirkepov current of ITS particle on the central (O+1)^3
//     nodes of the bin in REGISTERS, z-plane by z-plane (30 values per plane for order 2;
//     structural zeros are not stored).  The 32 lanes are then summed through a small per-warp
//     shared-memory transpose: lanes write [value][lane], pairs of lanes add up one row each
//     (128-bit conflict-free loads) and keep a RUNNING sum per (plane, value) in registers for all
//     iterations of the bin; when the bin is done one shared-memory atomic per value adds it to
//     the J tile.  This is the GPU analogue of the reference's sorted `reduce_add` path
//     (primitives.hpp:798-809): one scatter per bin instead of one per particle.
//   * the particle data of the NEXT iteration is prefetched with cp.async straight into a per-warp
//     staging buffer (no registers held across the iteration, nothing to spill), and the bins of a
//     tile are handed out dynamically (shared counter) so that the warps of a CTA finish together.
//   * particles that change bin ("movers", a few per cent) additionally touch nodes outside the
//     central mesh.  They leave a compact record (old and new position, bin) in a per-warp list; when
//     the list fills up, lanes = (mover, axis) expand ten records at a time into full 1-D weight
//     tables and lanes = (mover, face node) add the extra values to the J tile (9 nodes per
//     single-axis mover; the rare multi-axis mover walks its whole (O+3)^3 mesh).
#include "common.cuh"

#include <algorithm>

namespace nixb200
{
namespace
{
#ifndef NIX_PUSH_WARPS
#define NIX_PUSH_WARPS 6
#endif
#ifndef NIX_BATCH_CAS
#define NIX_BATCH_CAS 0 // movers: run the compare-and-swap loops of a lane's fp64 shared atomics in lock step
#endif
#ifndef NIX_PUSH_MINB
#define NIX_PUSH_MINB 2
#endif
constexpr int NWARP   = NIX_PUSH_WARPS;
constexpr int THREADS = 32 * NWARP;
constexpr int MAXMOV  = 24; // compact mover records per warp (old + new position, bin: 64 bytes each)
constexpr int CREC    = 8;  // doubles per compact record
constexpr int XGROUP  = 10; // movers expanded to full 1-D weight records at a time (3 axes x 10 = 30 lanes)
constexpr unsigned FULL = 0xffffffffu;
// per-warp reduction scratch: RROWS rows (one per plane value) of 2 x 16 lane slots, each half padded
// to 18 doubles: 64-bit stores of a half-warp and 128-bit loads of a quarter-warp are conflict-free
constexpr int RROWS = 16, RHALF = 18, RSTRIDE = 2 * RHALF;
constexpr int RED_DOUBLES_W = RROWS * RSTRIDE;
constexpr int PF_DOUBLES_W  = 2 * 6 * 32; // cp.async staging: 2 stages x 6 components x 32 lanes
#ifndef NIX_GATHER_UNROLL
#define NIX_GATHER_UNROLL 1 // 1: six copies of the gather body (no weight selects, larger code)
#endif

template <int O>
struct Cfg {
  static constexpr int NW = O + 2; // stencil width the staged E/B tile is padded by (interp.hpp)
  static constexpr int N1 = O + 1; // support of one shape function
  static constexpr int NS = O + 3; // deposit mesh width     (esirkepov.hpp:326-328)
  // values per z-plane of the central mesh kept in registers: rho, Jx (x slots 2..), Jy (y slots 2..),
  // Jz (planes 2..); the skipped entries are structural zeros for particles that stay in their bin
  static constexpr int P_RHO = 0;
  static constexpr int P_JX  = N1 * N1;
  static constexpr int P_JY  = P_JX + N1 * (N1 - 1);
  static constexpr int P_JZ  = P_JY + (N1 - 1) * N1;
  static constexpr int PV    = P_JZ + N1 * N1;         // 12 / 30 / 56
  static constexpr int NPASS = (PV + RROWS - 1) / RROWS; // reduction passes per plane
  // mover record (doubles): per axis (z,y,x): S0[NS] DS[NS] CP[NS]; then 2 doubles of ints
  static constexpr int REC = 9 * NS + 2;
  // bins per CTA (compile-time, so that every shared-memory offset of the gather is an immediate);
  // smaller chunks simply use part of the box
  static constexpr int TZ = 4, TY = 4, TX = 9;
  static constexpr int EZ = TZ + NW - 1, EY = TY + NW - 1, EX = TX + NW - 1; // staged E/B tile
  static constexpr int JZ = TZ + NS - 1, JY = TY + NS - 1, JX = TX + NS - 1; // J tile
  static constexpr int EB_DOUBLES  = (EZ * EY * EX * 6 + 15) / 16 * 16;
  // J tile, component-major: s_j[comp * JC + node] -- lanes that add the same component of neighbouring
  // nodes hit neighbouring banks (node-major [node][4] puts them 32 bytes apart: 4-way conflicts)
  static constexpr int JN = JZ * JY * JX, JC = (JN + 7) / 8 * 8 + 2;
  static constexpr int J_DOUBLES   = (4 * JC + 15) / 16 * 16;
  static constexpr int REC_DOUBLES = (NWARP * MAXMOV * CREC + 15) / 16 * 16;
};

struct SmemLayout {
  int    eb_doubles, j_doubles, rec_doubles, red_doubles, pf_doubles;
  size_t bytes;
};
// int region (offsets in ints)
constexpr int I_BAR  = 0;                               // mbarrier (2 ints)
constexpr int I_ANY  = 2;                               // tile has particles
constexpr int I_NEXT = 3;                               // next bin to hand out
constexpr int I_TBL  = 8;                               // [64]  J-tile offset of every plane value
constexpr int I_DCNT = I_TBL + 64;                      // [NWARP][32] leavers of the current bin per direction
constexpr int I_MLST = I_DCNT + NWARP * 32;             // [NWARP][MAXMOV] single-axis mover records
constexpr int I_CS   = (I_MLST + NWARP * MAXMOV + 1) / 2 * 2; // [TZ*TY][TX+1] first particle of every bin
constexpr int I_CG   = I_CS + 4 * 4 * (9 + 1);          // ChunkGeo (8-byte aligned)
constexpr int SMEM_INTS = (I_CG + (int)((sizeof(ChunkGeo) + 7) / 8 * 2) + 3) / 4 * 4;
static_assert(I_CG % 2 == 0, "s_cg must be 8-byte aligned");

template <int O>
__host__ __device__ inline SmemLayout smem_layout()
{
  using C = Cfg<O>;
  static_assert(C::TZ * C::TY * (C::TX + 1) <= 160 && C::PV <= 64, "int region sizes");
  SmemLayout L;
  L.eb_doubles  = C::EB_DOUBLES;
  L.j_doubles   = C::J_DOUBLES;
  L.rec_doubles = C::REC_DOUBLES;
  L.red_doubles = NWARP * RED_DOUBLES_W;
  L.pf_doubles  = NWARP * PF_DOUBLES_W;
  L.bytes       = sizeof(double) * ((size_t)L.eb_doubles + L.j_doubles + L.rec_doubles + L.red_doubles + L.pf_doubles) +
            SMEM_INTS * sizeof(int);
  return L;
}

// ---- mbarrier / TMA wrappers (inline PTX) -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0,
                                            int c1, int c2, int c3, int c4)
{
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ void cp_async8(void* dst, const void* src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- shape functions (primitives.hpp:257-298), association order preserved ----------------------
template <int O, bool S>
__device__ __forceinline__ void shape_mc(double x, double X, double rdx, double* s)
{
  double delta = mul<S>(sub<S>(x, X), rdx);
  if constexpr (O == 1) {
    s[0] = sub<S>(1.0, delta);
    s[1] = delta;
  } else if constexpr (O == 2) {
    double w0 = delta;
    double w1 = sub<S>(0.5, w0);
    double w2 = add<S>(0.5, w0);
    s[0]      = mul<S>(mul<S>(0.50, w1), w1);
    s[1]      = sub<S>(0.75, mul<S>(w0, w0));
    s[2]      = mul<S>(mul<S>(0.50, w2), w2);
  } else {
    const double a  = 1 / 6.0;
    double       w1 = delta;
    double       w2 = sub<S>(1.0, delta);
    double       w1_pow2 = mul<S>(w1, w1);
    double       w2_pow2 = mul<S>(w2, w2);
    double       w1_pow3 = mul<S>(w1_pow2, w1);
    double       w2_pow3 = mul<S>(w2_pow2, w2);
    s[0] = mul<S>(a, w2_pow3);
    s[1] = mul<S>(a, add<S>(sub<S>(4.0, mul<S>(6.0, w1_pow2)), mul<S>(3.0, w1_pow3)));
    s[2] = mul<S>(a, add<S>(sub<S>(4.0, mul<S>(6.0, w2_pow2)), mul<S>(3.0, w2_pow3)));
    s[3] = mul<S>(a, w1_pow3);
  }
}

// One field component gathered over its exact support (interp3d_impl_sorted, interp.hpp:95-113:
// same nesting and summation order; the reference's extra stencil slot carries a zero weight).
// e points at the first support node of the component; sy / sz = row / plane strides in doubles.
template <int O, bool S>
__device__ __forceinline__ double gather1(const double* __restrict__ e, const double* wz, const double* wy,
                                          const double* wx)
{
  constexpr int sy = Cfg<O>::EX * 6, sz = Cfg<O>::EY * Cfg<O>::EX * 6;
  double        rz = 0.0;
#pragma unroll
  for (int jz = 0; jz <= O; jz++) {
    double ry = 0.0;
#pragma unroll
    for (int jy = 0; jy <= O; jy++) {
      double        rx = 0.0;
      const double* p  = e + jz * sz + jy * sy;
#pragma unroll
      for (int jx = 0; jx <= O; jx++) rx = mad<S>(p[jx * 6], wx[jx], rx);
      ry = mad<S>(rx, wy[jy], ry);
    }
    rz = mad<S>(ry, wz[jz], rz);
  }
  return rz;
}

struct Kparams {
  Geo             geo;
  const ChunkGeo* cg;
  double*         uj;
  SpeciesDev      sp;
  double          delt, dt1, q;
  double          qdxdt[3]; // q * del/dt per axis (z,y,x)
  int*            err;
};

// 1-D deposit weights of one particle on the central slots 1..N1 of the (O+3) mesh, per axis (z,y,x):
// S0 = old weights, DS = new - old, CP[j] = sum of DS over slots < j (esirkepov.hpp:167-174)
template <int O>
struct Wts {
  double s0[3][O + 1], ds[3][O + 1], cp[3][O + 1];
};

// Esirkepov current of one particle on z-plane jz of the central mesh, added into acc[PV]:
//   rho[x] += q S1y S1z S1x[x]                                                   esirkepov.hpp:155-164
//   Jx[x]  += -q dx/dt ((S0y+DSy/2) S0z + (S0y/2+DSy/3) DSz) CPx[x]                        :177-195
//   Jy[x]  += -q dy/dt CPy ((S0z+DSz/2) S0x[x] + (S0z/2+DSz/3) DSx[x])                     :198-216
//   Jz[x]  += -q dz/dt CPz ((S0x+DSx/2)[x] S0y + (S0x/2+DSx/3)[x] DSy)                     :219-237
template <int O>
__device__ __forceinline__ void plane_accumulate(const double s0z, const double dsz, const double cpz,
                                                 const bool with_jz, const Wts<O>& w, const double q,
                                                 const double* qd, double* acc)
{
  using C          = Cfg<O>;
  constexpr int N1 = C::N1;
  const double  A = 1.0 / 2, B = 1.0 / 3;
  const double  qs1z = q * (s0z + dsz);
  const double  azy = -qd[1] * (s0z + A * dsz), bzy = -qd[1] * (A * s0z + B * dsz);
  const double  fz  = -qd[0] * cpz;
#pragma unroll
  for (int jy = 0; jy < N1; jy++) {
    const double s0y = w.s0[1][jy], dsy = w.ds[1][jy];
    const double ar  = qs1z * (s0y + dsy);
    const double wx  = -qd[2] * ((s0y + A * dsy) * s0z + (A * s0y + B * dsy) * dsz);
#pragma unroll
    for (int jx = 0; jx < N1; jx++) {
      const double s0x = w.s0[2][jx], dsx = w.ds[2][jx];
      acc[C::P_RHO + jy * N1 + jx] = fma(ar, s0x + dsx, acc[C::P_RHO + jy * N1 + jx]);
      if (jx >= 1) acc[C::P_JX + jy * (N1 - 1) + jx - 1] = fma(wx, w.cp[2][jx], acc[C::P_JX + jy * (N1 - 1) + jx - 1]);
      if (jy >= 1) {
        const double wy = azy * s0x + bzy * dsx;
        acc[C::P_JY + (jy - 1) * N1 + jx] = fma(wy, w.cp[1][jy], acc[C::P_JY + (jy - 1) * N1 + jx]);
      }
    }
  }
  if (with_jz) { // plane 0 carries no Jz of the register set (structural zero / mover extra)
#pragma unroll
    for (int jy = 0; jy < N1; jy++) {
      const double s0y = w.s0[1][jy], dsy = w.ds[1][jy];
#pragma unroll
      for (int jx = 0; jx < N1; jx++) {
        const double s0x = w.s0[2][jx], dsx = w.ds[2][jx];
        const double wz  = (s0x + A * dsx) * s0y + (A * s0x + B * dsx) * dsy;
        acc[C::P_JZ + jy * N1 + jx] = fma(fz, wz, acc[C::P_JZ + jy * N1 + jx]);
      }
    }
  }
}

// Sum the plane values acc[0..PV) of the 32 lanes through the warp's scratch and add them to the
// running sums of the bin: pass p handles the values [16p, 16p+16); lane l = (row l>>1, half l&1)
// adds up 16 of the 32 lane slots of its row, the two halves meet with one shuffle.  Afterwards
// BOTH lanes of a pair hold the sum of value 16p + (l>>1) in bsum[p].
template <int O>
__device__ __forceinline__ void plane_reduce(const double* acc, double* my_red, int lane, double* bsum)
{
  using C = Cfg<O>;
  const int col = (lane >> 4) * RHALF + (lane & 15);
  const double2* src = reinterpret_cast<const double2*>(my_red + (lane >> 1) * RSTRIDE + (lane & 1) * RHALF);
#pragma unroll
  for (int p = 0; p < C::NPASS; p++) {
#pragma unroll
    for (int r = 0; r < RROWS; r++)
      if (p * RROWS + r < C::PV) my_red[r * RSTRIDE + col] = acc[p * RROWS + r];
    __syncwarp();
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const double2 a = src[i], b = src[i + 1];
      s0 += a.x;
      s1 += a.y;
      s2 += b.x;
      s3 += b.y;
    }
    double sum = (s0 + s1) + (s2 + s3);
    sum += __shfl_xor_sync(FULL, sum, 1);
    bsum[p] += sum;
    __syncwarp();
  }
}

// N independent fp64 additions to shared memory.  fp64 shared-memory atomics are compare-and-swap
// loops; running the N loops of a lane in lock step overlaps their round trips.
template <int N>
__device__ __forceinline__ void atomic_add_batch(double* const* ad, const double* val, bool* todo)
{
#if !NIX_BATCH_CAS
#pragma unroll
  for (int k = 0; k < N; k++)
    if (todo[k]) atomicAdd(ad[k], val[k]);
#else
  unsigned long long old[N];
#pragma unroll
  for (int k = 0; k < N; k++) old[k] = todo[k] ? *reinterpret_cast<const volatile unsigned long long*>(ad[k]) : 0ull;
  bool any = true;
  while (any) {
    any = false;
#pragma unroll
    for (int k = 0; k < N; k++) {
      if (todo[k]) {
        const unsigned long long want = __double_as_longlong(__longlong_as_double(old[k]) + val[k]);
        const unsigned long long got  = atomicCAS(reinterpret_cast<unsigned long long*>(ad[k]), old[k], want);
        todo[k]                       = got != old[k];
        old[k]                        = got;
        any                           = any || todo[k];
      }
    }
  }
#endif
}

// Add the extra values (outside the register-resident set) of up to XGROUP expanded mover records
// (per axis z,y,x: S0[NS] DS[NS] CP[NS], then cellbase and mover code) to the J tile.
template <int O>
__device__ __forceinline__ void flush_group(double* s_j, const double* myrec, int* myml, int nrec, double q,
                                            double qdz, double qdy, double qdx)
{
  using C          = Cfg<O>;
  constexpr int N1 = C::N1, NS = C::NS, JY = C::JY, JX = C::JX;
  const int     lane = threadIdx.x & 31;
  __syncwarp();
  // single-axis movers: lanes = (record, face node (u,v)), N1*N1 nodes each
  int nsingle = 0;
  for (int m = 0; m < nrec; m++) {
    const int* ri = reinterpret_cast<const int*>(myrec + m * C::REC + 9 * NS);
    if (ri[1] >= 0) {
      if (lane == 0) myml[nsingle] = m;
      nsingle++;
    }
  }
  __syncwarp();
  const double A = 1.0 / 2, B = 1.0 / 3;
  auto         node = [&](const double* r, int jz, int jy, int jx, double& rho, double& wx, double& wy, double& wz) {
    const double s0z = r[0 * NS + jz], dsz = r[1 * NS + jz];
    const double s0y = r[3 * NS + jy], dsy = r[4 * NS + jy];
    const double s0x = r[6 * NS + jx], dsx = r[7 * NS + jx];
    rho = q * (s0z + dsz) * (s0y + dsy) * (s0x + dsx);
    wx  = -qdx * ((s0y + A * dsy) * s0z + (A * s0y + B * dsy) * dsz);
    wy  = -qdy * ((s0z + A * dsz) * s0x + (A * s0z + B * dsz) * dsx);
    wz  = -qdz * ((s0x + A * dsx) * s0y + (A * s0x + B * dsx) * dsy);
  };
  for (int it0 = 0; it0 < nsingle * N1 * N1; it0 += 32) {
    const int gi = it0 + lane;
    if (gi < nsingle * N1 * N1) {
      const int     m  = myml[gi / (N1 * N1)];
      const int     uv = gi % (N1 * N1);
      const double* r  = myrec + m * C::REC;
      const int*    ri = reinterpret_cast<const int*>(r + 9 * NS);
      const int     cbase = ri[0], ax = ri[1] >> 1, o = (ri[1] & 1) ? NS - 1 : 0;
      // mesh slots z,y,x: the moving axis sits on its outer slot, the two in-plane axes (ascending
      // axis order) run over the central slots
      const int u = uv / N1 + 1, v = uv % N1 + 1;
      const int jz = (ax == 0) ? o : u;
      const int jy = (ax == 1) ? o : ((ax == 0) ? u : v);
      const int jx = (ax == 2) ? o : v;
      double    rho, wx, wy, wz;
      node(r, jz, jy, jx, rho, wx, wy, wz);
      double*      dst = s_j + cbase + (jz * JY + jy) * JX + jx;
      const double vx = wx * r[8 * NS + jx], vy = wy * r[5 * NS + jy], vz = wz * r[2 * NS + jz];
      // low-side mover: the current through the first central face (slot 1) is carried by DS[0]
      const int     st = (ax == 0) ? JY * JX : ((ax == 1) ? JX : 1);
      const double  w1 = (ax == 0) ? wz * r[2 * NS + 1] : ((ax == 1) ? wy * r[5 * NS + 1] : wx * r[8 * NS + 1]);
      double* const ad[5]   = {dst, dst + C::JC, dst + 2 * C::JC, dst + 3 * C::JC, dst + st + (3 - ax) * C::JC};
      const double  val[5]  = {rho, vx, vy, vz, w1};
      bool          todo[5] = {rho != 0.0, vx != 0.0, vy != 0.0, vz != 0.0, o == 0 && w1 != 0.0};
      atomic_add_batch<5>(ad, val, todo);
    }
  }
  // multi-axis movers: the whole (O+3)^3 mesh minus what the register path already holds
  for (int m = 0; m < nrec; m++) {
    const double* r  = myrec + m * C::REC;
    const int*    ri = reinterpret_cast<const int*>(r + 9 * NS);
    if (ri[1] >= 0) continue;
    const int cbase = ri[0];
    for (int n = lane; n < NS * NS * NS; n += 32) {
      const int  jz = n / (NS * NS), jy = (n / NS) % NS, jx = n % NS;
      const bool central = jz >= 1 && jz <= N1 && jy >= 1 && jy <= N1 && jx >= 1 && jx <= N1;
      double     rho, wx, wy, wz;
      node(r, jz, jy, jx, rho, wx, wy, wz);
      double*      dst = s_j + cbase + (jz * JY + jy) * JX + jx;
      const double vx = wx * r[8 * NS + jx], vy = wy * r[5 * NS + jy], vz = wz * r[2 * NS + jz];
      if (!central && rho != 0.0) atomicAdd(dst, rho);
      if (!(central && jx >= 2) && vx != 0.0) atomicAdd(dst + C::JC, vx);
      if (!(central && jy >= 2) && vy != 0.0) atomicAdd(dst + 2 * C::JC, vy);
      if (!(central && jz >= 2) && vz != 0.0) atomicAdd(dst + 3 * C::JC, vz);
    }
  }
  __syncwarp();
}


struct MoverGeo { // what the expansion needs of the geometry, by value (no param-space pointers)
  double del[3], rdel[3];
  int    is_odd;
};

// Flush the compact mover records of one warp: XGROUP at a time, lanes = (mover, axis) recompute the
// 1-D deposit weights over the whole (O+3) mesh exactly as the main loop does for the central slots
// (ss[0][.][1..O+1] = old weights, ss[1][.][1+shift..] = new weights, test_esirkepov.cpp:1060-1085;
// DS and its running sum CP, esirkepov.hpp:167-174) into the warp's reduction scratch, then
// flush_group adds the values outside the register-resident set.  Kept out of line: it runs once per
// ~20 movers.
template <int O, bool S>
__device__ __noinline__ void flush_movers(double* s_j, const double* crec, double* xrec, int* myml, int nrec,
                                          const ChunkGeo* c, MoverGeo mg, double q, double qdz, double qdy, double qdx)
{
  using C          = Cfg<O>;
  constexpr int N1 = C::N1, NS = C::NS;
  static_assert(XGROUP * C::REC <= RED_DOUBLES_W && XGROUP * 3 <= 32, "expanded records live in the reduction scratch");
  const int lane = threadIdx.x & 31;
  for (int m0 = 0; m0 < nrec; m0 += XGROUP) {
    const int ng = min(XGROUP, nrec - m0);
    __syncwarp();
    if (lane < 3 * ng) {
      const int     m = lane / 3, a = lane - 3 * m;
      const double* cr = crec + (size_t)(m0 + m) * CREC;
      const int*    ci = reinterpret_cast<const int*>(cr + 6);
      const double  xo = cr[a], xn = cr[3 + a];
      const int     bin = (a == 0) ? ci[2] : ((a == 1) ? (ci[3] & 0xffff) : (ci[3] >> 16));
      const int     ki = bin - mg.is_odd;
      const int     k1 = digitize(xn, c->off[a], mg.rdel[a]) - mg.is_odd;
      const int     sft = k1 - ki;
      double        wi[N1], wn[N1];
      shape_mc<O, S>(xo, add<S>(c->imin[a], mul<S>((double)ki, mg.del[a])), mg.rdel[a], wi);
      shape_mc<O, S>(xn, add<S>(c->imin[a], mul<S>((double)k1, mg.del[a])), mg.rdel[a], wn);
      double* r  = xrec + m * C::REC + 3 * a * NS;
      double  cp = 0.0;
#pragma unroll
      for (int j = 0; j < NS; j++) {
        const double s0 = (j >= 1 && j <= O + 1) ? wi[j - 1] : 0.0;
        const double vm = (j >= 0 && j <= O) ? wn[j] : 0.0;         // shift -1: slot j <- wn[j]
        const double v0 = (j >= 1 && j <= O + 1) ? wn[j - 1] : 0.0; // shift  0
        const double vp = (j >= 2 && j <= O + 2) ? wn[j - 2] : 0.0; // shift +1
        const double s1 = (sft == 0) ? v0 : ((sft < 0) ? vm : vp);
        const double ds = s1 - s0;
        r[j]            = s0;
        r[NS + j]       = ds;
        r[2 * NS + j]   = cp;
        cp += ds;
      }
      if (a == 0) {
        int* ri = reinterpret_cast<int*>(xrec + m * C::REC + 9 * NS);
        ri[0]   = ci[0];
        ri[1]   = ci[1];
      }
    }
    __syncwarp();
    flush_group<O>(s_j, xrec, myml, ng, q, qdz, qdy, qdx);
  }
  __syncwarp();
}

template <int O, bool S>
__global__ void __launch_bounds__(THREADS, NIX_PUSH_MINB) k_push_deposit(const __grid_constant__ CUtensorMap tmap,
                                                             const Kparams P)
{
  using C          = Cfg<O>;
  constexpr int N1 = C::N1;
  constexpr int NS = C::NS;
  constexpr int PV = C::PV;
  const Geo&    g    = P.geo;
  const int     tid  = threadIdx.x;
  const int     lane = tid & 31;
  const int     warp = tid >> 5;

  // ---- decode the work item -------------------------------------------------------------------
  const int tl = blockIdx.x % g.ntile;
  const int ch = blockIdx.x / g.ntile;
  int       b0[3], nbn[3];
  {
    int t[3] = {tl / (g.ntl[1] * g.ntl[2]), (tl / g.ntl[2]) % g.ntl[1], tl % g.ntl[2]};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      b0[a]  = t[a] * g.tile[a];
      nbn[a] = min(g.tile[a], g.nc[a] - b0[a]);
    }
  }
  constexpr int EY = C::EY, EX = C::EX;
  constexpr int JZ = C::JZ, JY = C::JY, JX = C::JX;
  constexpr int esy = EX * 6, esz = EY * EX * 6;

  extern __shared__ __align__(1024) double smem_d[];
  const SmemLayout L = smem_layout<O>();
  double*   s_eb   = smem_d;
  double*   s_j    = smem_d + L.eb_doubles;
  double*   s_rec  = s_j + L.j_doubles;
  double*   s_red  = s_rec + L.rec_doubles;
  double*   s_pf   = s_red + L.red_doubles;
  int*      s_int  = reinterpret_cast<int*>(s_pf + L.pf_doubles);
  uint64_t* s_bar  = reinterpret_cast<uint64_t*>(s_int + I_BAR);
  int*      s_any  = s_int + I_ANY;
  int*      s_next = s_int + I_NEXT;
  int*      s_tbl  = s_int + I_TBL;
  int*      s_dcnt = s_int + I_DCNT;
  int*      s_mlst = s_int + I_MLST;
  ChunkGeo* s_cg   = reinterpret_cast<ChunkGeo*>(s_int + I_CG);
  int*      s_cs   = s_int + I_CS;

  const int32_t* __restrict__ start = P.sp.start;
  const int cellkey0 = ch * g.ncell;

  // ---- particle ranges of the tile's bins -> shared memory; empty tile? (e.g. the rounding-guard
  //      bin layer of even orders) ------------------------------------------------------------------
  constexpr int CSW = C::TX + 1; // row width of s_cs
  if (tid == 0) {
    *s_any  = 0;
    *s_next = NWARP; // the first NWARP bins are taken by the warps directly
  }
  for (int t = tid; t < (int)(sizeof(ChunkGeo) / sizeof(int)); t += THREADS)
    reinterpret_cast<int*>(s_cg)[t] = reinterpret_cast<const int*>(P.cg + ch)[t];
  for (int t = tid; t < nbn[0] * nbn[1] * CSW; t += THREADS) {
    const int r = t / CSW, x = t % CSW;
    if (x <= nbn[2]) {
      const int rz = b0[0] + r / nbn[1], ry = b0[1] + r % nbn[1];
      s_cs[t] = start[(cellkey0 + (rz * g.R[1] + ry) * g.R[2] + b0[2] + x) * LANES];
    }
  }
  __syncthreads();
  if (tid < nbn[0] * nbn[1] && s_cs[tid * CSW + nbn[2]] != s_cs[tid * CSW]) atomicOr(s_any, 1);
  __syncthreads();
  if (*s_any == 0) return;

  // tile origins (array indices)
  const int Lb  = g.nb;
  const int ez0 = b0[0] - g.is_odd - g.half + Lb, ey0 = b0[1] - g.is_odd - g.half + Lb,
            ex0 = b0[2] - g.is_odd - g.half + Lb;
  const int jz0 = ez0 - 1, jy0 = ey0 - 1, jx0 = ex0 - 1;

  if (tid == 0) {
    mbar_init(s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(s_bar, (uint32_t)(C::EZ * EY * EX * 6 * sizeof(double)));
    tma_load_5d(s_eb, &tmap, s_bar, 0, ex0, ey0, ez0, ch);
  }
  for (int t = tid; t < L.j_doubles; t += THREADS) s_j[t] = 0.0;
  for (int t = tid; t < L.red_doubles; t += THREADS) s_red[t] = 0.0;
  for (int v = tid; v < PV; v += THREADS) {
    // plane value v -> (component, jy, jx) in mesh slots (central index + 1)
    int comp, jy, jx;
    if (v < C::P_JX) {
      comp = 0, jy = v / N1, jx = v % N1;
    } else if (v < C::P_JY) {
      int r = v - C::P_JX;
      comp = 1, jy = r / (N1 - 1), jx = r % (N1 - 1) + 1;
    } else if (v < C::P_JZ) {
      int r = v - C::P_JY;
      comp = 2, jy = r / N1 + 1, jx = r % N1;
    } else {
      int r = v - C::P_JZ;
      comp = 3, jy = r / N1, jx = r % N1;
    }
    s_tbl[v] = comp * C::JC + (jy + 1) * JX + (jx + 1);
  }
  __syncthreads();

  const ChunkGeo& c   = *s_cg; // staged in shared memory: read again and again by every iteration
  const int       cb  = P.sp.cbase[ch];
  const size_t    cap = P.sp.cap;
  double* __restrict__ xu = P.sp.xu;
  double* myrec = s_rec + (size_t)warp * MAXMOV * CREC;
  MoverGeo mgeo;
#pragma unroll
  for (int a = 0; a < 3; a++) mgeo.del[a] = g.del[a], mgeo.rdel[a] = g.rdel[a];
  mgeo.is_odd = g.is_odd;
  int*    mydc  = s_dcnt + warp * 32;
  int*    myml  = s_mlst + warp * MAXMOV;

  mbar_wait(s_bar, 0);

  double* my_red = s_red + warp * RED_DOUBLES_W;
  double* my_pf  = s_pf + warp * PF_DOUBLES_W;
  constexpr int NPASS = C::NPASS;

  // running sums of the current bin: bsum[z][p] = value 16p + (lane>>1) of plane z
  double bsum[N1][NPASS];
#pragma unroll
  for (int z = 0; z < N1; z++)
#pragma unroll
    for (int p = 0; p < NPASS; p++) bsum[z][p] = 0.0;
  // add the running sums of a finished bin to the J tile (one lane of each pair per value)
  auto flush_bin = [&](int cellbase) {
#pragma unroll
    for (int z = 0; z < N1; z++)
#pragma unroll
      for (int p = 0; p < NPASS; p++) {
        const int    v   = p * RROWS + (lane >> 1);
        const double val = bsum[z][p];
        if ((lane & 1) == (p & 1) && v < PV && val != 0.0) // both lanes of a pair hold the sum: share the passes
          atomicAdd(s_j + cellbase + (z + 1) * JY * JX + s_tbl[v], val);
        bsum[z][p] = 0.0;
      }
  };

  // ---- bins of the tile, handed out dynamically; one iteration = up to 32 particles of one bin.
  //      The loop is flat and software-pipelined: the particle data of the NEXT iteration is
  //      requested (cp.async into the warp's staging buffer) before the current one is processed.
  const int ncell_t = nbn[0] * nbn[1] * nbn[2];
  // state of an iteration: bin cl = r * nbn[2] + lx (r = row lz * nbn[1] + ly), particles [i0, pe)
  auto advance = [&](int& cl, int& r, int& lx, int& i0, int& pe) -> bool {
    i0 += 32;
    while (i0 >= pe) { // next non-empty bin
      int nx = 0;
      if (lane == 0) nx = atomicAdd(s_next, 1);
      cl = __shfl_sync(FULL, nx, 0);
      if (cl >= ncell_t) return false;
      r  = cl / nbn[2];
      lx = cl - r * nbn[2];
      i0 = s_cs[r * CSW + lx];
      pe = s_cs[r * CSW + lx + 1];
    }
    return true;
  };
  auto prefetch = [&](int stage, int i) {
#pragma unroll
    for (int k = 0; k < 6; k++) cp_async8(my_pf + (stage * 6 + k) * 32 + lane, xu + soa(k, cap, i));
  };

  int  ncl = warp, nr = 0, nlx = 0, ni0 = 0, npe = 0;
  bool have = ncl < ncell_t;
  if (have) {
    nr   = ncl / nbn[2];
    nlx  = ncl - nr * nbn[2];
    ni0  = s_cs[nr * CSW + nlx] - 32;
    npe  = s_cs[nr * CSW + nlx + 1];
    have = advance(ncl, nr, nlx, ni0, npe);
  }
  int stage = 0;
  if (have && ni0 + lane < npe) prefetch(0, ni0 + lane);
  cp_async_commit();

  int  prev_cl = -1, nrec = 0;
  int  bz = 0, by = 0, bx = 0, cellbase = 0;
  bool had_leav = false; // warp-uniform
  const double* ecell = s_eb;

  while (have) {
    const int cl = ncl, cr = nr, clx = nlx, i0 = ni0, pe = npe;
    have = advance(ncl, nr, nlx, ni0, npe);
    cp_async_wait_all();
    double cur[6];
#pragma unroll
    for (int k = 0; k < 6; k++) cur[k] = my_pf[(stage * 6 + k) * 32 + lane];
    stage ^= 1;
    if (have && ni0 + lane < npe) prefetch(stage, ni0 + lane);
    cp_async_commit();

    if (cl != prev_cl) { // first iteration of a bin
      if (prev_cl >= 0) flush_bin(cellbase);
      const int lx = clx, lz = (nbn[1] == C::TY) ? cr / C::TY : cr / nbn[1], ly = cr - lz * nbn[1];
      bz = b0[0] + lz, by = b0[1] + ly, bx = b0[2] + lx;
      cellbase = (lz * JY + ly) * JX + lx; // J-tile node of mesh slot (0,0,0)
      ecell    = s_eb + (size_t)((lz * EY + ly) * EX + lx) * 6;
      had_leav = false;
      prev_cl  = cl;
      __syncwarp();
      if (lane < 27) mydc[lane] = 0;
      __syncwarp();
    }
    {
      const int  i     = i0 + lane;
      const bool valid = i < pe;
      int        dir   = 13;
      bool       dep_ok = false, mover = false;
      int        mcode = -1; // single-axis mover: axis*2 + (1 if high side); -1: multi-axis
      Wts<O>     w;
      double     mvn[3] = {0.0, 0.0, 0.0}; // new position (z,y,x), for the mover record

      // =============================== push ===============================
      if (valid) {
        const double pos[3] = {cur[2], cur[1], cur[0]}; // index 0,1,2 = z,y,x
        const double u[3]   = {cur[5], cur[4], cur[3]};

        int    ki[3], bh[3];
        double wi[3][N1], wh[3][N1];
        bool   sorted_ok = true;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          int ii = digitize(pos[a], c.off[a], g.rdel[a]);
          ki[a]  = ii - g.is_odd;
          int hh = digitize(pos[a], c.hoff[a], g.rdel[a]);
          shape_mc<O, S>(pos[a], add<S>(c.imin[a], mul<S>((double)ki[a], g.del[a])), g.rdel[a], wi[a]);
          shape_mc<O, S>(pos[a], add<S>(c.lo[a], mul<S>((double)hh, g.del[a])), g.rdel[a], wh[a]);
          // interp::shift_weights<O>(hh - ki, wh)  interp.hpp:154-160: the half-grid support starts
          // one node later; kept as a base offset instead of moving the weights
          bh[a]     = (hh - ki[a] > 0) ? 1 : 0;
          sorted_ok = sorted_ok && (ii == ((a == 0) ? bz : ((a == 1) ? by : bx)));
        }
        if (!sorted_ok) atomicOr(P.err, NIXB200_ERR_UNSORTED);

        // ---- gather: Ex Ey Ez Bx By Bz; half-grid axes: Ex x | Ey y | Ez z | Bx y,z | By x,z | Bz x,y
        // one loop body for the six components (code size: the loop must stay instruction-cache
        // resident); the results shift through f6 so that no register array is indexed dynamically
        double f6[6];
#pragma unroll
        for (int k = 0; k < 6; k++) f6[k] = 0.0;
#if NIX_GATHER_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int k = 0; k < 6; k++) {
          const bool hz = (0x1C >> k) & 1, hy = (0x2A >> k) & 1, hx = (0x31 >> k) & 1;
          double     wz[N1], wy[N1], wx[N1];
#pragma unroll
          for (int j = 0; j < N1; j++) {
            wz[j] = hz ? wh[0][j] : wi[0][j];
            wy[j] = hy ? wh[1][j] : wi[1][j];
            wx[j] = hx ? wh[2][j] : wi[2][j];
          }
          const double* e = ecell + (hz ? bh[0] : 0) * esz + (hy ? bh[1] : 0) * esy + (hx ? bh[2] : 0) * 6 + k;
          const double  f = gather1<O, S>(e, wz, wy, wx);
#pragma unroll
          for (int q = 0; q < 5; q++) f6[q] = f6[q + 1];
          f6[5] = f;
        }
        double ex = mul<S>(f6[0], P.dt1), ey = mul<S>(f6[1], P.dt1), ez = mul<S>(f6[2], P.dt1);
        double bxx = mul<S>(f6[3], P.dt1), byy = mul<S>(f6[4], P.dt1), bzz = mul<S>(f6[5], P.dt1);

        // ---- push_boris (primitives.hpp:165-189) ----------------------------------------------
        double ux = u[2], uy = u[1], uz = u[0];
        ux = add<S>(ux, ex);
        uy = add<S>(uy, ey);
        uz = add<S>(uz, ez);
        double gm = div_<S>(1.0, sqrt_<S>(add<S>(add<S>(add<S>(mul<S>(g.cc, g.cc), mul<S>(ux, ux)), mul<S>(uy, uy)), mul<S>(uz, uz))));
        bxx = mul<S>(bxx, gm);
        byy = mul<S>(byy, gm);
        bzz = mul<S>(bzz, gm);
        double bb = div_<S>(2.0, add<S>(add<S>(add<S>(1.0, mul<S>(bxx, bxx)), mul<S>(byy, byy)), mul<S>(bzz, bzz)));
        double vx = add<S>(ux, sub<S>(mul<S>(uy, bzz), mul<S>(uz, byy)));
        double vy = add<S>(uy, sub<S>(mul<S>(uz, bxx), mul<S>(ux, bzz)));
        double vz = add<S>(uz, sub<S>(mul<S>(ux, byy), mul<S>(uy, bxx)));
        ux = add<S>(ux, add<S>(mul<S>(sub<S>(mul<S>(vy, bzz), mul<S>(vz, byy)), bb), ex));
        uy = add<S>(uy, add<S>(mul<S>(sub<S>(mul<S>(vz, bxx), mul<S>(vx, bzz)), bb), ey));
        uz = add<S>(uz, add<S>(mul<S>(sub<S>(mul<S>(vx, byy), mul<S>(vy, bxx)), bb), ez));

        // ---- position update (lorentz_factor, primitives.hpp:158-161) ------------------------
        double uu  = add<S>(add<S>(mul<S>(ux, ux), mul<S>(uy, uy)), mul<S>(uz, uz));
        double gam = sqrt_<S>(add<S>(1.0, mul<S>(mul<S>(uu, g.rc), g.rc)));
        double dtg = div_<S>(P.delt, gam);
        double pn[3];
        pn[2] = add<S>(pos[2], mul<S>(ux, dtg));
        pn[1] = add<S>(pos[1], mul<S>(uy, dtg));
        pn[0] = add<S>(pos[0], mul<S>(uz, dtg));

        mvn[0] = pn[0], mvn[1] = pn[1], mvn[2] = pn[2];
        xu[soa(0, cap, i)] = pn[2];
        xu[soa(1, cap, i)] = pn[1];
        xu[soa(2, cap, i)] = pn[0];
        xu[soa(3, cap, i)] = ux;
        xu[soa(4, cap, i)] = uy;
        xu[soa(5, cap, i)] = uz;

        // ---- bin of the new position: count / classify ------------------------------------------
        int  i1[3], sft[3];
        bool cfl_ok = true;
        int  dcode  = 0, nmove = 0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          i1[a]  = digitize(pn[a], c.off[a], g.rdel[a]);
          int dd = (pn[a] >= c.hi[a]) - (pn[a] < c.lo[a]) + 1;
          dcode  = dcode * 3 + dd;
          sft[a] = (i1[a] - g.is_odd) - ki[a];
          cfl_ok = cfl_ok && (sft[a] >= -1) && (sft[a] <= 1);
          if (sft[a] != 0) {
            nmove++;
            mcode = a * 2 + (sft[a] > 0 ? 1 : 0);
          }
        }
        dir            = dcode;
        const int lnid = (i - cb) & (LANES - 1);
        if (dir == 13) {
          int key     = (cellkey0 + (i1[0] * g.R[1] + i1[1]) * g.R[2] + i1[2]) * LANES + lnid;
          P.sp.key[i] = key;
          atomicAdd(&P.sp.hist[key], 1);
        } else {
          P.sp.key[i] = -1;
          atomicAdd(&P.sp.oob[ch * LANES + lnid], 1);
        }
        if (!cfl_ok) atomicOr(P.err, NIXB200_ERR_CFL);
        dep_ok = cfl_ok && sorted_ok;
        mover  = dep_ok && nmove > 0;
        if (nmove > 1) mcode = -1;

        // ---- 1-D deposit weights -------------------------------------------------------------------
        // ss[0][.][1..O+1] = old weights; ss[1][.][1+sft..] = new weights (test_esirkepov.cpp:1060-1085)
#pragma unroll
        for (int a = 0; a < 3; a++) {
          double wn[N1];
          int    k1 = i1[a] - g.is_odd;
          shape_mc<O, S>(pn[a], add<S>(c.imin[a], mul<S>((double)k1, g.del[a])), g.rdel[a], wn);
          double cp = 0.0;
#pragma unroll
          for (int j = 0; j < NS; j++) {
            double s0 = (j >= 1 && j <= O + 1) ? wi[a][j - 1] : 0.0;
            double vm = (j >= 0 && j <= O) ? wn[j] : 0.0;         // sft = -1 : slot j <- wn[j]
            double v0 = (j >= 1 && j <= O + 1) ? wn[j - 1] : 0.0; // sft =  0
            double vp = (j >= 2 && j <= O + 2) ? wn[j - 2] : 0.0; // sft = +1
            double s1 = (sft[a] == 0) ? v0 : ((sft[a] < 0) ? vm : vp);
            if (!dep_ok) {
              s0 = 0.0;
              s1 = 0.0;
            }
            const double ds = s1 - s0; // ds3d, esirkepov.hpp:167-174
            if (j >= 1 && j <= N1) {
              w.s0[a][j - 1] = s0;
              w.ds[a][j - 1] = ds;
              w.cp[a][j - 1] = cp;
            }
            cp += ds;
          }
        }
      } else {
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int j = 0; j < N1; j++) w.s0[a][j] = w.ds[a][j] = w.cp[a][j] = 0.0;
      }

      // ---- leavers: ordered rank inside this bin per direction (warp shuffles) -----------------------
      const bool     leaver = valid && dir != 13;
      const unsigned lm     = __ballot_sync(FULL, leaver);
      if (lm) {
        had_leav = true;
        if (leaver) {
          const unsigned grpm = __match_any_sync(lm, dir);
          const int      rk   = __popc(grpm & ((1u << lane) - 1));
          const int      r    = mydc[dir] + rk;
          const int      bb3[3] = {bz, by, bx};
          const int      se   = slab_entry(g, dir, bb3);
          if (se < 0) atomicOr(P.err, NIXB200_ERR_CFL);
          int slot = atomicAdd(P.sp.nleave, 1);
          if (slot < P.sp.lcap && se >= 0) P.sp.lrec[slot] = make_int4(i, ch, se, (r << 5) | dir);
          else atomicOr(P.err, NIXB200_ERR_CAPACITY);
          __syncwarp(lm);
          if (rk == 0) mydc[dir] += __popc(grpm);
        }
        __syncwarp();
      }

      // =============================== deposit ===============================
      // plane by plane: evaluate on the lane's registers, sum the 32 lanes through the scratch
#pragma unroll
      for (int z = 0; z < N1; z++) {
        double acc[PV];
#pragma unroll
        for (int v = 0; v < PV; v++) acc[v] = 0.0;
        plane_accumulate<O>(w.s0[0][z], w.ds[0][z], w.cp[0][z], z >= 1, w, P.q, P.qdxdt, acc);
        plane_reduce<O>(acc, my_red, lane, bsum[z]);
      }

      // ---- movers: compact record (old / new position, bin); the list is flushed when it is full ----
      unsigned mm = __ballot_sync(FULL, mover);
      while (mm) {
        const int  room = MAXMOV - nrec;
        const int  rk   = __popc(mm & ((1u << lane) - 1));
        const bool take = mover && ((mm >> lane) & 1u) && rk < room;
        if (take) {
          double* r = myrec + (nrec + rk) * CREC;
          r[0] = cur[2], r[1] = cur[1], r[2] = cur[0]; // old z, y, x
          r[3] = mvn[0], r[4] = mvn[1], r[5] = mvn[2]; // new z, y, x
          int* ri = reinterpret_cast<int*>(r + 6);
          ri[0]   = cellbase;
          ri[1]   = mcode;
          ri[2]   = bz;
          ri[3]   = by | (bx << 16);
        }
        const unsigned taken = __ballot_sync(FULL, take);
        nrec += __popc(taken);
        mm &= ~taken;
        if (mm || nrec > MAXMOV - 8) { // full (or nearly: the next iteration brings a few more)
          flush_movers<O, S>(s_j, myrec, my_red, myml, nrec, s_cg, mgeo, P.q, P.qdxdt[0], P.qdxdt[1], P.qdxdt[2]);
          nrec = 0;
        }
      }
    }

    // last iteration of a bin: its leavers per direction -> slab counts (scanned by k_mig_scan)
    if (had_leav && (!have || ncl != cl)) {
      __syncwarp();
      if (lane < 27 && lane != 13 && mydc[lane] > 0) {
        const int bb3[3] = {bz, by, bx};
        const int se     = slab_entry(g, lane, bb3);
        if (se >= 0) P.sp.slabcnt[(size_t)ch * g.slaboff[27] + se] = mydc[lane];
      }
    }
  }
  if (prev_cl >= 0) flush_bin(cellbase);
  if (nrec) flush_movers<O, S>(s_j, myrec, my_red, myml, nrec, s_cg, mgeo, P.q, P.qdxdt[0], P.qdxdt[1], P.qdxdt[2]);

  // ---- flush the J tile: the CTA's single scatter to global memory --------------------------------
  __syncthreads();
  double* __restrict__ ujc = P.uj + (size_t)ch * g.M[0] * g.M[1] * g.M[2] * 4;
  for (int t = tid; t < JZ * JY * JX * 4; t += THREADS) {
    const double v = s_j[(t & 3) * C::JC + (t >> 2)];
    if (v != 0.0) {
      const int k = t & 3, n = t >> 2;
      const int gx = jx0 + n % JX, gy = jy0 + (n / JX) % JY, gz = jz0 + n / (JX * JY);
      if (gx >= 0 && gx < g.M[2] && gy >= 0 && gy < g.M[1] && gz >= 0 && gz < g.M[0])
        atomicAdd(&ujc[(((size_t)gz * g.M[1] + gy) * g.M[2] + gx) * 4 + k], v);
    }
  }
}

template <int O, bool S>
int launch_t(const PushArgs& a, const CUtensorMap* tmap, cudaStream_t st)
{
  Kparams P;
  P.geo  = a.geo;
  P.cg   = a.cg;
  P.uj   = a.uj;
  P.sp   = a.sp;
  P.delt = a.delt;
  P.dt1  = 0.5 * a.sp.q / a.sp.m * a.delt; // ref_driver.cpp: dt1
  P.q    = a.sp.q;
  for (int d = 0; d < 3; d++) P.qdxdt[d] = a.sp.q * (a.geo.del[d] / a.delt);
  P.err = a.err;
  size_t smem = smem_layout<O>().bytes;
  static bool attr_set = false;
  if (!attr_set) {
    NIX_CUDA(cudaFuncSetAttribute(k_push_deposit<O, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    attr_set = true;
  }
  int nblocks = a.geo.nchunk * a.geo.ntile;
  k_push_deposit<O, S><<<nblocks, THREADS, smem, st>>>(*tmap, P);
  NIX_LAUNCHED();
  return 0;
}
} // namespace

size_t push_smem_bytes(const Geo& g)
{
  switch (g.order) {
  case 1: return smem_layout<1>().bytes;
  case 2: return smem_layout<2>().bytes;
  default: return smem_layout<3>().bytes;
  }
}

void push_tile_box(int order, int& tz, int& ty, int& tx)
{
  (void)order;
  tz = Cfg<2>::TZ;
  ty = Cfg<2>::TY;
  tx = Cfg<2>::TX;
}

// The bins-per-CTA box is a compile-time constant of the kernel (Cfg<O>::TZ/TY/TX); small chunks use
// a part of it, long rows are cut into several tiles.
int choose_push_tile(Geo& g)
{
  const int box[3] = {Cfg<2>::TZ, Cfg<2>::TY, Cfg<2>::TX};
  static_assert(Cfg<1>::TZ == Cfg<2>::TZ && Cfg<3>::TZ == Cfg<2>::TZ && Cfg<1>::TY == Cfg<2>::TY &&
                    Cfg<3>::TY == Cfg<2>::TY && Cfg<1>::TX == Cfg<2>::TX && Cfg<3>::TX == Cfg<2>::TX,
                "one tile box for all orders");
  g.ntile = 1;
  for (int a = 0; a < 3; a++) {
    int n     = (g.nc[a] + box[a] - 1) / box[a];      // tiles along this axis
    g.tile[a] = (g.nc[a] + n - 1) / n;                // balanced, <= box
    g.ntl[a]  = (g.nc[a] + g.tile[a] - 1) / g.tile[a];
    g.ntile *= g.ntl[a];
  }
  return 0;
}

int launch_push_deposit(const PushArgs& a, const CUtensorMap* tmap, bool strict, cudaStream_t st)
{
  // leaver bookkeeping of this step
  NIX_CUDA(cudaMemsetAsync(a.sp.slabcnt, 0, sizeof(int32_t) * (size_t)a.geo.nchunk * a.geo.slaboff[27], st));
  NIX_CUDA(cudaMemsetAsync(a.sp.oob, 0, sizeof(int32_t) * a.geo.nchunk * LANES, st));
  NIX_CUDA(cudaMemsetAsync(a.sp.nleave, 0, sizeof(int32_t), st));
  switch (a.geo.order) {
  case 1: return strict ? launch_t<1, true>(a, tmap, st) : launch_t<1, false>(a, tmap, st);
  case 2: return strict ? launch_t<2, true>(a, tmap, st) : launch_t<2, false>(a, tmap, st);
  case 3: return strict ? launch_t<3, true>(a, tmap, st) : launch_t<3, false>(a, tmap, st);
  default: set_error("order must be 1, 2 or 3"); return 1;
  }
}
} // namespace nixb200
