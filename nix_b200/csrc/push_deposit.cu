// push_deposit.cu -- gather + Boris push + position update + count (k_push) and the Esirkepov
// deposit (k_deposit), two kernels per species and step (sm_100a)
//
// Replace, per particle (composition: oracle/ref/ref_driver.cpp):
//   digitize / shape_mc<O>      primitives.hpp:46-58,257-298,519-532
//   interp::shift_weights<O>    interp.hpp:149-172
//   interp::interp3d<O>         interp.hpp:95-113,217-230
//   push_boris / push_vay / push_higuera_cary, lorentz_factor    primitives.hpp:158-253
//   esirkepov::deposit3d<O>     esirkepov.hpp:155-237,326-340
//   append_current3d<O>         primitives.hpp:778-834
//   XtensorParticle::count      xtensor_particle.hpp:324-357   (of the NEW positions)
//   XtensorHaloParticle3D::pre_pack classification  xtensor_halo3d.hpp:288-302
//
// Work item of both kernels = one (tz x ty x tx) tile of bins of one chunk = one CTA.
//
// k_push (8 warps, 2 CTAs/SM):
//   * TMA (cp.async.bulk.tensor.5d) stages the E/B tile of the CTA -- ghosts included -- into shared
//     memory; an mbarrier signals arrival.
//   * a WARP owns one x-row of bins at a time (rows are handed out dynamically).  The container is
//     cell-sorted, so a row is one contiguous particle range: the warp streams through it 32
//     consecutive particles at a time (coalesced SoA loads through a two-stage cp.async staging
//     buffer, coalesced stores), every lane addressing the E/B tile from ITS bin.
//   * lane = particle: weights, gather over the EXACT (O+1)^3 support of every component, momentum
//     update, move; the old position goes to xv[0:3] (the reference's temporary array,
//     test_esirkepov.cpp:1046-1051), the new state in place; bin of the new position -> key +
//     histogram, or a leaver record with its ordered rank inside (bin, direction).
//
// k_deposit (4 warps, 3 CTAs/SM): a J tile of the CTA's deposit footprint lives in shared memory and
// is flushed to global memory ONCE per CTA with red.global.add.f64.
//   * a WARP owns one bin at a time (bins are handed out dynamically); per iteration its lanes load
//     old and new position of 32 particles and leave the 1-D deposit weights (per axis S0, DS and the
//     running sum of DS) in a per-warp scratch.
//   * then the lanes change roles: lane = (particle slot, z-plane) -- 8 slots x 4 lanes for orders 2
//     and 3, 16 x 2 for order 1 -- evaluates, round by round, the Esirkepov current of one particle on
//     ONE z-plane of the central (O+1)^3 nodes of the bin and adds it to a register-resident
//     accumulator of that plane (30 values for order 2; structural zeros are not stored) that lives
//     for ALL particles of the bin.  The lanes of a slot read the same scratch words, so the
//     64/128-bit loads of a round are broadcasts.
//   * when the bin is done the accumulators of the slots are summed through the idle scratch
//     (conflict-free strided loads) and one shared-memory atomic per value adds the bin's current to
//     the J tile: the GPU analogue of the reference's sorted `reduce_add` path
//     (primitives.hpp:798-809) with one scatter per BIN instead of one per particle, and no
//     cross-lane reduction per iteration.
//   * particles that change bin ("movers", a few per cent) additionally touch nodes outside the
//     central mesh.  They leave a compact record (old and new position, bin) in a per-warp list; when
//     the list is full, lanes = (mover, axis) expand ten records at a time into full 1-D weight
//     tables and lanes = (mover, face node) add the extra values to the J tile (9 nodes per
//     single-axis mover; the rare multi-axis mover walks its whole (O+3)^3 mesh).
#include "common.cuh"

#include <algorithm>

namespace nixb200
{
namespace
{
#ifndef NIX_MAXMOV
#define NIX_MAXMOV 20
#endif
#ifndef NIX_MOVER_AGGREGATE
#define NIX_MOVER_AGGREGATE 0 // 1: sum the lanes of a mover flush that target the same nodes before the atomics
                               // (match_any + shuffles; measured slower: k_deposit 23.3 instead of 20.8 ms per step)
#endif
#ifndef NIX_J_FIXED
#define NIX_J_FIXED 0 // 1 (EXPERIMENT, never run on a device: prepared after the round's GPU budget was spent): the fp64 J
                      // tile of k_deposit holds 64-bit FIXED-POINT words and is added to with two NATIVE 32-bit shared-memory
                      // atomics (low word, carry into the high word: exact and order-independent) instead of the
                      // compare-and-swap loops an fp64 shared-memory atomicAdd compiles to (profiles/r02f_lines_dep_*.md:
                      // 25.8 % of the stall samples of the electron launch); converted back once per CTA in the global
                      // flush.  Resolution: 2^-45 of the largest per-particle contribution.  fp32 instantiations unchanged.
#endif
#ifndef NIX_MOV_FLUSH
#define NIX_MOV_FLUSH NIX_MAXMOV // flush whole groups of XGROUP records: the expansion and face-node lanes stay busy
#endif
constexpr int MAXMOV  = NIX_MAXMOV; // compact mover records per warp (old + new position, bin)
template <typename T>
constexpr int crec_len() { return 6 + 4 * (int)sizeof(int) / (int)sizeof(T); } // reals per compact record: old + new position, 4 ints
#define CREC (crec_len<T>())
constexpr int XGROUP  = 10; // movers expanded to full 1-D weight records at a time (3 axes x 10 = 30 lanes)
constexpr unsigned FULL = 0xffffffffu;
constexpr int PF_DOUBLES_W = 6 * 32; // one cp.async staging stage: 6 components x 32 lanes (a lane fills its own slots)
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
  // mover record (doubles): per axis (z,y,x): S0[NS] DS[NS] CP[NS]; then 2 doubles of ints
  static constexpr int REC = 9 * NS + 2;
  // deposit rounds: lane = (particle slot, z-plane); PLW lanes per particle, NPS particles per round
  static constexpr int PLW = (N1 <= 2) ? 2 : 4;
  static constexpr int NPS = 32 / PLW;
  // per-warp weight scratch (doubles).  Axis record of y and x: S0[N1] DS[N1] CP[1..N1-1] as NPR
  // 16-byte pairs, [axis][pair][particle][2]; z: [plane][particle][2] = (S0z, DSz); [plane-1][particle] = CPz
  static constexpr int NPR   = (3 * N1) / 2;
  // (plane stride ZP = 33 particles: the lanes of a slot read different planes of the same particle,
  // which a stride of 32 would put into the same banks)
  static constexpr int ZP    = 33;
  static constexpr int WQ_Z  = 2 * NPR * 64;
  static constexpr int WQ_ZC = WQ_Z + N1 * 2 * ZP;
  static constexpr int WQ_DOUBLES = WQ_ZC + (N1 - 1) * ZP;
  // the scratch also holds the expanded mover records of flush_movers
  static constexpr int SCR = ((WQ_DOUBLES > XGROUP * REC ? WQ_DOUBLES : XGROUP * REC) + 15) / 16 * 16;
  // bins per CTA (compile-time, so that every shared-memory offset of the gather is an immediate);
  // smaller chunks simply use part of the box
  static constexpr int TZ = 4, TY = 4, TX = 9;
  // staged E/B tile: the stencil box plus dummy rows / columns (the TMA box is simply larger) chosen so
  // that the row and plane strides, modulo the 128 bytes of the banks, keep the four half-grid base
  // offsets of every component in different bank pairs (unpadded, both strides are 64 B for order 2)
#ifndef NIX_TILE_PAD
#define NIX_TILE_PAD 1
#endif
  static constexpr int PADX = !NIX_TILE_PAD ? 0 : (O == 2 ? 1 : 0), PADY = !NIX_TILE_PAD ? 0 : (O == 2 ? 3 : (O == 3 ? 2 : 0));
  static constexpr int EZ = TZ + NW - 1, EY = TY + NW - 1 + PADY, EX = TX + NW - 1 + PADX;
  static constexpr int JZ = TZ + NS - 1, JY = TY + NS - 1, JX = TX + NS - 1; // J tile
  template <typename T>
  static constexpr int eb_len() { return (EZ * EY * EX * field_stride<T>() + 15) / 16 * 16; } // staged E/B tile, in T
  // J tile, component-major: s_j[comp * JC + node] -- lanes that add the same component of neighbouring
  // nodes hit neighbouring banks (node-major [node][4] puts them 32 bytes apart: 4-way conflicts)
  static constexpr int JN = JZ * JY * JX, JC = (JN + 7) / 8 * 8 + 2;
  static constexpr int J_DOUBLES   = (4 * JC + 15) / 16 * 16;
};

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

template <typename T>
__device__ __forceinline__ void cp_async_elem(T* dst, const T* src)
{
  if constexpr (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
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
template <int O, bool S, typename T>
__device__ __forceinline__ void shape_mc(T x, T X, T rdx, T* s)
{
  T delta = mul<S>(sub<S>(x, X), rdx);
  if constexpr (O == 1) {
    s[0] = sub<S>(T(1.0), delta);
    s[1] = delta;
  } else if constexpr (O == 2) {
    T w0 = delta;
    T w1 = sub<S>(T(0.5), w0);
    T w2 = add<S>(T(0.5), w0);
    s[0]      = mul<S>(mul<S>(T(0.50), w1), w1);
    s[1]      = sub<S>(T(0.75), mul<S>(w0, w0));
    s[2]      = mul<S>(mul<S>(T(0.50), w2), w2);
  } else {
    const T a  = 1 / T(6.0);
    T       w1 = delta;
    T       w2 = sub<S>(T(1.0), delta);
    T       w1_pow2 = mul<S>(w1, w1);
    T       w2_pow2 = mul<S>(w2, w2);
    T       w1_pow3 = mul<S>(w1_pow2, w1);
    T       w2_pow3 = mul<S>(w2_pow2, w2);
    s[0] = mul<S>(a, w2_pow3);
    s[1] = mul<S>(a, add<S>(sub<S>(T(4.0), mul<S>(T(6.0), w1_pow2)), mul<S>(T(3.0), w1_pow3)));
    s[2] = mul<S>(a, add<S>(sub<S>(T(4.0), mul<S>(T(6.0), w2_pow2)), mul<S>(T(3.0), w2_pow3)));
    s[3] = mul<S>(a, w1_pow3);
  }
}

// One field component gathered over its exact support (interp3d_impl_sorted, interp.hpp:95-113:
// same nesting and summation order; the reference's extra stencil slot carries a zero weight).
// e points at the first support node of the component; sy / sz = row / plane strides in doubles.
template <int O, bool S, typename T>
__device__ __forceinline__ T gather1(const T* __restrict__ e, const T* wz, const T* wy,
                                          const T* wx)
{
  constexpr int FC = field_stride<T>(), sy = Cfg<O>::EX * FC, sz = Cfg<O>::EY * Cfg<O>::EX * FC;
  T        rz = T(0.0);
#pragma unroll
  for (int jz = 0; jz <= O; jz++) {
    T ry = T(0.0);
#pragma unroll
    for (int jy = 0; jy <= O; jy++) {
      T        rx = T(0.0);
      const T* p  = e + jz * sz + jy * sy;
#pragma unroll
      for (int jx = 0; jx <= O; jx++) rx = mad<S>(p[jx * FC], wx[jx], rx);
      ry = mad<S>(rx, wy[jy], ry);
    }
    rz = mad<S>(ry, wz[jz], rz);
  }
  return rz;
}

// kernel parameters; every real-valued constant is already in the kernel's real type T, so that no
// fp64 operand sneaks into an fp32 instantiation
template <typename T>
struct KparamsT {
  Geo             geo;
  const ChunkGeo* cg;
  T*              uj;
  T*              xu; // [NCT][cap]
  T*              xv;
  SpeciesDev      sp; // (integer tables; its xu / xv are untyped bytes)
  T               delt, dt1, q;
  T               qdxdt[3]; // q * del/dt per axis (z,y,x)
  T               del[3], rdel[3], cc, rc;
  int*            err;
  int             pusher;   // NIXB200_PUSH_*
};

// 1-D deposit weights of one particle on the central slots 1..N1 of the (O+3) mesh, per axis (z,y,x):
// S0 = old weights, DS = new - old, CP[j] = sum of DS over slots < j (esirkepov.hpp:167-174)
template <int O, typename T>
struct Wts {
  T s0[3][O + 1], ds[3][O + 1], cp[3][O + 1];
};

// Esirkepov current of one particle on ONE z-plane of the central mesh, added into acc[PV]:
//   rho[x] += q S1y S1z S1x[x]                                                   esirkepov.hpp:155-164
//   Jx[x]  += -q dx/dt ((S0y+DSy/2) S0z + (S0y/2+DSy/3) DSz) CPx[x]                        :177-195
//   Jy[x]  += -q dy/dt CPy ((S0z+DSz/2) S0x[x] + (S0z/2+DSz/3) DSx[x])                     :198-216
//   Jz[x]  += -q dz/dt CPz ((S0x+DSx/2)[x] S0y + (S0x/2+DSx/3)[x] DSy)                     :219-237
// ty / tx = axis records of y / x: S0[N1] DS[N1] CP[1..N1-1].  cpz is zero on plane 0 (structural
// zero of the register set; the first face of a low-side mover is added by the mover path) and all
// three z weights are zero on an idle lane, which then adds exact zeros.
template <int O, typename T>
__device__ __forceinline__ void plane_accumulate(const T s0z, const T dsz, const T cpz,
                                                 const T* ty, const T* tx, const T q,
                                                 const T* qd, T* acc)
{
  using C          = Cfg<O>;
  constexpr int N1 = C::N1;
  const T  A = T(1.0) / 2, B = T(1.0) / 3;
  const T  qs1z = q * (s0z + dsz);
  const T  azy = -qd[1] * (s0z + A * dsz), bzy = -qd[1] * (A * s0z + B * dsz);
  const T  fz  = -qd[0] * cpz;
#pragma unroll
  for (int jy = 0; jy < N1; jy++) {
    const T s0y = ty[jy], dsy = ty[N1 + jy];
    const T ar  = qs1z * (s0y + dsy);
    const T wx  = -qd[2] * ((s0y + A * dsy) * s0z + (A * s0y + B * dsy) * dsz);
#pragma unroll
    for (int jx = 0; jx < N1; jx++) {
      const T s0x = tx[jx], dsx = tx[N1 + jx];
      acc[C::P_RHO + jy * N1 + jx] = fma(ar, s0x + dsx, acc[C::P_RHO + jy * N1 + jx]);
      if (jx >= 1)
        acc[C::P_JX + jy * (N1 - 1) + jx - 1] = fma(wx, tx[2 * N1 + jx - 1], acc[C::P_JX + jy * (N1 - 1) + jx - 1]);
      if (jy >= 1) {
        const T wy = azy * s0x + bzy * dsx;
        acc[C::P_JY + (jy - 1) * N1 + jx] = fma(wy, ty[2 * N1 + jy - 1], acc[C::P_JY + (jy - 1) * N1 + jx]);
      }
      const T wz = (s0x + A * dsx) * s0y + (A * s0x + B * dsx) * dsy;
      acc[C::P_JZ + jy * N1 + jx] = fma(fz, wz, acc[C::P_JZ + jy * N1 + jx]);
    }
  }
}

// One deposit round: lane = (slot ps, plane pl) adds plane pl of particle p = first + ps to acc.
template <int O, typename T>
__device__ __forceinline__ void deposit_round(const T* wq, int p, int plc, bool pl_on, const T q,
                                              const T* qd, T* acc)
{
  using C           = Cfg<O>;
  constexpr int NPR = C::NPR;
  using R2 = typename Real<T>::vec2;
  const R2* q2 = reinterpret_cast<const R2*>(wq);
  T         ty[2 * NPR], tx[2 * NPR];
#pragma unroll
  for (int pr = 0; pr < NPR; pr++) {
    const R2 a = q2[pr * 32 + p], b = q2[(NPR + pr) * 32 + p];
    ty[2 * pr] = a.x, ty[2 * pr + 1] = a.y;
    tx[2 * pr] = b.x, tx[2 * pr + 1] = b.y;
  }
  const R2 z2 = q2[C::WQ_Z / 2 + plc * C::ZP + p];
  const T  zc = wq[C::WQ_ZC + (plc >= 1 ? plc - 1 : 0) * C::ZP + p];
  const T  s0z = pl_on ? z2.x : T(0.0), dsz = pl_on ? z2.y : T(0.0), cpz = (pl_on && plc >= 1) ? zc : T(0.0);
  plane_accumulate<O, T>(s0z, dsz, cpz, ty, tx, q, qd, acc);
}

#if NIX_J_FIXED
#include "jfixed.cuh"
#endif

// N independent fp64 additions to shared memory.  fp64 shared-memory atomics are compare-and-swap
// loops; running the N loops of a lane in lock step overlaps their round trips.
template <int N, typename T>
__device__ __forceinline__ void atomic_add_batch(T* const* ad, const T* val, bool* todo)
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
template <int O, typename T>
__device__ __forceinline__ void flush_group(T* s_j, const T* myrec, int* myml, int nrec, T q,
                                            T qdz, T qdy, T qdx)
{
  using C          = Cfg<O>;
  constexpr int N1 = C::N1, NS = C::NS, JY = C::JY, JX = C::JX;
  const int     lane = threadIdx.x & 31;
  __syncwarp();
  // single-axis movers: lanes = (record, face node (u,v)), N1*N1 nodes each
  const int      mycode = (lane < nrec) ? reinterpret_cast<const int*>(myrec + lane * C::REC + 9 * NS)[1] : 0;
  const unsigned single = __ballot_sync(FULL, lane < nrec && mycode >= 0);
  unsigned       multi  = __ballot_sync(FULL, lane < nrec && mycode < 0);
  const int      nsingle = __popc(single);
  if (lane < nrec && mycode >= 0) myml[__popc(single & ((1u << lane) - 1))] = lane;
  __syncwarp();
  const T A = T(1.0) / 2, B = T(1.0) / 3;
  auto         node = [&](const T* r, int jz, int jy, int jx, T& rho, T& wx, T& wy, T& wz) {
    const T s0z = r[0 * NS + jz], dsz = r[1 * NS + jz];
    const T s0y = r[3 * NS + jy], dsy = r[4 * NS + jy];
    const T s0x = r[6 * NS + jx], dsx = r[7 * NS + jx];
    rho = q * (s0z + dsz) * (s0y + dsy) * (s0x + dsx);
    wx  = -qdx * ((s0y + A * dsy) * s0z + (A * s0y + B * dsy) * dsz);
    wy  = -qdy * ((s0z + A * dsz) * s0x + (A * s0z + B * dsz) * dsx);
    wz  = -qdz * ((s0x + A * dsx) * s0y + (A * s0x + B * dsx) * dsy);
  };
  for (int it0 = 0; it0 < nsingle * N1 * N1; it0 += 32) {
    const int  gi     = it0 + lane;
    const bool active = gi < nsingle * N1 * N1;
    T*         dst    = s_j;
    T          val[5] = {T(0.0), T(0.0), T(0.0), T(0.0), T(0.0)};
    int        ax = 0, key = -1 - lane; // (a key of its own for an idle lane)
    bool       low = false;
    if (active) {
      const int     m  = myml[gi / (N1 * N1)];
      const int     uv = gi % (N1 * N1);
      const T* r  = myrec + m * C::REC;
      const int*    ri = reinterpret_cast<const int*>(r + 9 * NS);
      const int     cbase = ri[0], o = (ri[1] & 1) ? NS - 1 : 0;
      ax = ri[1] >> 1;
      // mesh slots z,y,x: the moving axis sits on its outer slot, the two in-plane axes (ascending
      // axis order) run over the central slots
      const int u = uv / N1 + 1, v = uv % N1 + 1;
      const int jz = (ax == 0) ? o : u;
      const int jy = (ax == 1) ? o : ((ax == 0) ? u : v);
      const int jx = (ax == 2) ? o : v;
      T    rho, wx, wy, wz;
      node(r, jz, jy, jx, rho, wx, wy, wz);
      const int nd = cbase + (jz * JY + jy) * JX + jx;
      dst = s_j + nd;
      low = o == 0;
      val[0] = rho, val[1] = wx * r[8 * NS + jx], val[2] = wy * r[5 * NS + jy], val[3] = wz * r[2 * NS + jz];
      // low-side mover: the current through the first central face (slot 1) is carried by DS[0]
      val[4] = !low ? T(0.0) : ((ax == 0) ? wz * r[2 * NS + 1] : ((ax == 1) ? wy * r[5 * NS + 1] : wx * r[8 * NS + 1]));
      key    = (nd * 4 + ax) * 2 + (low ? 1 : 0);
    }
#if NIX_MOVER_AGGREGATE
    // Movers of one bin that leave through the same face add to the SAME nodes: compare-and-swap loops on one
    // address serialise (profiles/r02f: 207 wavefronts where 49 would do).  Lanes with the same target are
    // summed with shuffles first and one of them issues the atomics.
    {
      const unsigned peers  = __match_any_sync(FULL, key);
      const int      cnt    = __popc(peers), leader = __ffs(peers) - 1;
      const int      maxcnt = (int)__reduce_max_sync(FULL, (unsigned)cnt);
      for (int k = 1; k < maxcnt; k++) {
        const int src = (k < cnt) ? (int)__fns(peers, 0, k + 1) : lane;
#pragma unroll
        for (int c = 0; c < 5; c++) {
          const T t = __shfl_sync(FULL, val[c], src);
          if (lane == leader && k < cnt) val[c] += t;
        }
      }
      if (lane != leader) {
#pragma unroll
        for (int c = 0; c < 5; c++) val[c] = T(0.0);
      }
    }
#endif
    if (active) {
      const int st = (ax == 0) ? JY * JX : ((ax == 1) ? JX : 1);
      T* const  ad[5]   = {dst, dst + C::JC, dst + 2 * C::JC, dst + 3 * C::JC, dst + st + (3 - ax) * C::JC};
      bool      todo[5] = {val[0] != T(0.0), val[1] != T(0.0), val[2] != T(0.0), val[3] != T(0.0), low && val[4] != T(0.0)};
#if NIX_J_FIXED
      const T jsc = jt_scale<T>(q, qdz, qdy, qdx);
#pragma unroll
      for (int k = 0; k < 5; k++)
        if (todo[k]) jt_add<T>(ad[k], val[k], jsc);
#else
      atomic_add_batch<5, T>(ad, val, todo);
#endif
    }
  }
  // multi-axis movers: the whole (O+3)^3 mesh minus what the register path already holds
  while (multi) {
    const int m = __ffs(multi) - 1;
    multi &= multi - 1;
    const T* r  = myrec + m * C::REC;
    const int*    ri = reinterpret_cast<const int*>(r + 9 * NS);
    const int cbase = ri[0];
    for (int n = lane; n < NS * NS * NS; n += 32) {
      const int  jz = n / (NS * NS), jy = (n / NS) % NS, jx = n % NS;
      const bool central = jz >= 1 && jz <= N1 && jy >= 1 && jy <= N1 && jx >= 1 && jx <= N1;
      T     rho, wx, wy, wz;
      node(r, jz, jy, jx, rho, wx, wy, wz);
      T*      dst = s_j + cbase + (jz * JY + jy) * JX + jx;
      const T vx = wx * r[8 * NS + jx], vy = wy * r[5 * NS + jy], vz = wz * r[2 * NS + jz];
#if NIX_J_FIXED
      const T jsc = jt_scale<T>(q, qdz, qdy, qdx);
      if (!central && rho != T(0.0)) jt_add<T>(dst, rho, jsc);
      if (!(central && jx >= 2) && vx != T(0.0)) jt_add<T>(dst + C::JC, vx, jsc);
      if (!(central && jy >= 2) && vy != T(0.0)) jt_add<T>(dst + 2 * C::JC, vy, jsc);
      if (!(central && jz >= 2) && vz != T(0.0)) jt_add<T>(dst + 3 * C::JC, vz, jsc);
#else
      if (!central && rho != T(0.0)) atomicAdd(dst, rho);
      if (!(central && jx >= 2) && vx != T(0.0)) atomicAdd(dst + C::JC, vx);
      if (!(central && jy >= 2) && vy != T(0.0)) atomicAdd(dst + 2 * C::JC, vy);
      if (!(central && jz >= 2) && vz != T(0.0)) atomicAdd(dst + 3 * C::JC, vz);
#endif
    }
  }
  __syncwarp();
}


template <typename T>
struct MoverGeoT { // what the expansion needs of the geometry, by value (no param-space pointers)
  T   del[3], rdel[3];
  int is_odd;
};

// Flush the compact mover records of one warp: XGROUP at a time, lanes = (mover, axis) recompute the
// 1-D deposit weights over the whole (O+3) mesh exactly as the main loop does for the central slots
// (ss[0][.][1..O+1] = old weights, ss[1][.][1+shift..] = new weights, test_esirkepov.cpp:1060-1085;
// DS and its running sum CP, esirkepov.hpp:167-174) into the warp's reduction scratch, then
// flush_group adds the values outside the register-resident set.  Kept out of line: it runs once per
// ~20 movers.
template <int O, bool S, typename T>
__device__ __noinline__ void flush_movers(T* s_j, const T* crec, T* xrec, int* myml, int nrec,
                                          const ChunkGeoT<T>* c, MoverGeoT<T> mg, T q, T qdz, T qdy, T qdx)
{
  using C          = Cfg<O>;
  constexpr int N1 = C::N1, NS = C::NS;
  static_assert(XGROUP * C::REC <= C::SCR && XGROUP * 3 <= 32, "expanded records live in the weight scratch");
  const int lane = threadIdx.x & 31;
  for (int m0 = 0; m0 < nrec; m0 += XGROUP) {
    const int ng = min(XGROUP, nrec - m0);
    __syncwarp();
    if (lane < 3 * ng) {
      const int     m = lane / 3, a = lane - 3 * m;
      const T* cr = crec + (size_t)(m0 + m) * CREC;
      const int*    ci = reinterpret_cast<const int*>(cr + 6);
      const T  xo = cr[a], xn = cr[3 + a];
      const int     bin = (a == 0) ? ci[2] : ((a == 1) ? (ci[3] & 0xffff) : (ci[3] >> 16));
      const int     ki = bin - mg.is_odd;
      const int     k1 = digitize(xn, c->off[a], mg.rdel[a]) - mg.is_odd;
      const int     sft = k1 - ki;
      T        wi[N1], wn[N1];
      shape_mc<O, S, T>(xo, add<S>(c->imin[a], mul<S>((T)ki, mg.del[a])), mg.rdel[a], wi);
      shape_mc<O, S, T>(xn, add<S>(c->imin[a], mul<S>((T)k1, mg.del[a])), mg.rdel[a], wn);
      T* r  = xrec + m * C::REC + 3 * a * NS;
      T  cp = T(0.0);
#pragma unroll
      for (int j = 0; j < NS; j++) {
        const T s0 = (j >= 1 && j <= O + 1) ? wi[j - 1] : T(0.0);
        const T vm = (j >= 0 && j <= O) ? wn[j] : T(0.0);         // shift -1: slot j <- wn[j]
        const T v0 = (j >= 1 && j <= O + 1) ? wn[j - 1] : T(0.0); // shift  0
        const T vp = (j >= 2 && j <= O + 2) ? wn[j - 2] : T(0.0); // shift +1
        const T s1 = (sft == 0) ? v0 : ((sft < 0) ? vm : vp);
        const T ds = s1 - s0;
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
    flush_group<O, T>(s_j, xrec, myml, ng, q, qdz, qdy, qdx);
  }
  __syncwarp();
}

// =================================================================================================
// Split path: k_push (gather + Boris + move + count + leaver classification) and k_deposit (Esirkepov
// deposit from the old and new positions).  k_push keeps the reference's own data flow -- old
// position to xv[0:3], new state in place (test_esirkepov.cpp:1046-1051) -- and needs neither the
// deposit accumulators nor bin-aligned iterations: a warp streams through a whole x-row of bins, 32
// consecutive particles at a time, every lane addressing the E/B tile from ITS bin.  k_deposit is
// the bin-owned part (lane roles, accumulators, movers) fed by the 48 bytes of positions only.
// =================================================================================================
#ifndef NIX_P_WARPS
#define NIX_P_WARPS 8
#endif
#ifndef NIX_P_MINB
#define NIX_P_MINB 2
#endif
#ifndef NIX_D_WARPS
#define NIX_D_WARPS 4
#endif
#ifndef NIX_D_MINB
#define NIX_D_MINB 3
#endif
constexpr int PWARPS = NIX_P_WARPS, PTHREADS = 32 * PWARPS;
constexpr int DWARPS = NIX_D_WARPS, DTHREADS = 32 * DWARPS;

// int regions (offsets in ints)
constexpr int IP_BAR = 0, IP_ANY = 2, IP_NEXT = 3;
constexpr int IP_DCNT = 8;                          // [PWARPS][32] carried leaver counts of the open bin
constexpr int IP_CS   = IP_DCNT + PWARPS * 32;      // [TZ*TY][TX+1]
constexpr int IP_CG   = IP_CS + 4 * 4 * (9 + 1);
constexpr int IP_INTS = (IP_CG + (int)((sizeof(ChunkGeo) + 7) / 8 * 2) + 3) / 4 * 4;
static_assert(IP_CG % 2 == 0, "s_cg must be 8-byte aligned");
constexpr int ID_ANY = 0, ID_NEXT = 1;
constexpr int ID_TBL  = 8;                          // [64]
constexpr int ID_MLST = ID_TBL + 64;                // [DWARPS][MAXMOV]
constexpr int ID_CS   = (ID_MLST + DWARPS * MAXMOV + 1) / 2 * 2;
constexpr int ID_CG   = ID_CS + 4 * 4 * (9 + 1);
constexpr int ID_INTS = (ID_CG + (int)((sizeof(ChunkGeo) + 7) / 8 * 2) + 3) / 4 * 4;
static_assert(ID_CG % 2 == 0, "s_cg must be 8-byte aligned");

template <int O, typename T>
__host__ __device__ inline size_t push_smem()
{
  return sizeof(T) * ((size_t)Cfg<O>::template eb_len<T>() + PWARPS * 2 * PF_DOUBLES_W) + IP_INTS * sizeof(int);
}
template <int O, typename T>
__host__ __device__ inline size_t deposit_smem()
{
  using C = Cfg<O>;
  return sizeof(T) * ((size_t)C::J_DOUBLES + (DWARPS * MAXMOV * CREC + 15) / 16 * 16 + DWARPS * C::SCR +
                           DWARPS * PF_DOUBLES_W) + ID_INTS * sizeof(int);
}

template <int O, bool S, typename T>
__global__ void __launch_bounds__(PTHREADS, NIX_P_MINB) k_push(const __grid_constant__ CUtensorMap tmap, const KparamsT<T> P)
{
  using C          = Cfg<O>;
  constexpr int N1 = C::N1;
  const Geo&    g    = P.geo;
  const int     tid  = threadIdx.x;
  const int     lane = tid & 31;
  const int     warp = tid >> 5;

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
  constexpr int FC = field_stride<T>(), esy = EX * FC, esz = EY * EX * FC;

  if (P.err[1]) return; // a particle store overflowed earlier: the cell ranges no longer describe memory
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  T* smem_d = reinterpret_cast<T*>(smem_raw);
  T*   s_eb   = smem_d;
  T*   s_pf   = smem_d + C::template eb_len<T>();
  int*      s_int  = reinterpret_cast<int*>(s_pf + PWARPS * 2 * PF_DOUBLES_W); // two staging stages
  uint64_t* s_bar  = reinterpret_cast<uint64_t*>(s_int + IP_BAR);
  int*      s_any  = s_int + IP_ANY;
  int*      s_next = s_int + IP_NEXT;
  int*      s_dcnt = s_int + IP_DCNT;
  int*      s_cs   = s_int + IP_CS;
  ChunkGeoT<T>* s_cg = reinterpret_cast<ChunkGeoT<T>*>(s_int + IP_CG);

  const int32_t* __restrict__ start = P.sp.start;
  const int cellkey0 = ch * g.ncell;
  constexpr int CSW = C::TX + 1;
  if (tid == 0) {
    *s_any  = 0;
    *s_next = PWARPS; // the first PWARPS rows are taken by the warps directly
  }
  stage_chunk_geo(s_cg, P.cg + ch, tid, PTHREADS);
  for (int t = tid; t < nbn[0] * nbn[1] * CSW; t += PTHREADS) {
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

  const int Lb  = g.nb;
  const int ez0 = b0[0] - g.is_odd - g.half + Lb, ey0 = b0[1] - g.is_odd - g.half + Lb,
            ex0 = b0[2] - g.is_odd - g.half + Lb;
  if (tid == 0) {
    mbar_init(s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(s_bar, (uint32_t)(C::EZ * EY * EX * FC * sizeof(T)));
    tma_load_5d(s_eb, &tmap, s_bar, 0, ex0, ey0, ez0, ch);
  }

  const ChunkGeoT<T>& c = *s_cg;
  const int       cb  = P.sp.cbase[ch];
  const size_t    cap = P.sp.cap;
  T* __restrict__ xu = P.xu;
  T* __restrict__ xv = P.xv;
  int*          mydc = s_dcnt + warp * 32;
  T*       my_pf = s_pf + warp * 2 * PF_DOUBLES_W;
  const int     nrow  = nbn[0] * nbn[1];

  // rows of the tile, handed out dynamically; one iteration = 32 consecutive particles of the row
  auto advance = [&](int& r, int& i0, int& pe) -> bool {
    i0 += 32;
    while (i0 >= pe) {
      int nx = 0;
      if (lane == 0) nx = atomicAdd(s_next, 1);
      r = __shfl_sync(FULL, nx, 0);
      if (r >= nrow) return false;
      i0 = s_cs[r * CSW];
      pe = s_cs[r * CSW + nbn[2]];
    }
    return true;
  };
  auto prefetch = [&](int stage, int i) {
#pragma unroll
    for (int k = 0; k < 6; k++) cp_async_elem(my_pf + (stage * 6 + k) * 32 + lane, xu + soa(k, cap, i));
  };

  int  nr = warp, ni0 = 0, npe = 0;
  bool have = nr < nrow;
  if (have) {
    ni0  = s_cs[nr * CSW] - 32;
    npe  = s_cs[nr * CSW + nbn[2]];
    have = advance(nr, ni0, npe);
  }
  int stage = 0;
  if (have && ni0 + lane < npe) prefetch(0, ni0 + lane);
  cp_async_commit();

  mbar_wait(s_bar, 0);

  int  prev_r = -1, open_lx = -1;
  bool carry = false; // mydc may hold counts of the open bin
  int  bz = 0, by = 0;
  const T* erow = s_eb;
  const int*    csrow = s_cs;

  while (have) {
    const int r = nr, i0 = ni0, pe = npe;
    have = advance(nr, ni0, npe);
    cp_async_wait_all();
    // two staging stages: the next iteration's particles are requested before this one's are used
    const T* mine = my_pf + stage * 6 * 32 + lane;
    stage ^= 1;
    if (have && ni0 + lane < npe) prefetch(stage, ni0 + lane);
    cp_async_commit();
    if (r != prev_r) {
      const int lz = (nbn[1] == C::TY) ? r / C::TY : r / nbn[1], ly = r - lz * nbn[1];
      bz = b0[0] + lz, by = b0[1] + ly;
      erow    = s_eb + (size_t)((lz * EY + ly) * EX) * FC;
      csrow   = s_cs + r * CSW;
      open_lx = -1;
      carry   = false;
      prev_r  = r;
    }
    const int  i     = i0 + lane;
    const bool valid = i < pe;
    int        dir   = 13;
    int        lxc   = 0;

    if (valid) {
      const T pos[3] = {mine[2 * 32], mine[1 * 32], mine[0]}; // index 0,1,2 = z,y,x
      int    ki[3], bh[3], ii[3];
      T wi[3][N1], wh[3][N1];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        ii[a]  = digitize(pos[a], c.off[a], P.rdel[a]);
        ki[a]  = ii[a] - g.is_odd;
        int hh = digitize(pos[a], c.hoff[a], P.rdel[a]);
        shape_mc<O, S, T>(pos[a], add<S>(c.imin[a], mul<S>((T)ki[a], P.del[a])), P.rdel[a], wi[a]);
        shape_mc<O, S, T>(pos[a], add<S>(c.lo[a], mul<S>((T)hh, P.del[a])), P.rdel[a], wh[a]);
        bh[a] = (hh - ki[a] > 0) ? 1 : 0;
      }
      // the particle must sit in the bin that owns its slot of the cell-sorted container
      const int lx = ii[2] - b0[2];
      lxc          = min(max(lx, 0), nbn[2] - 1);
      const bool sorted_ok = ii[0] == bz && ii[1] == by && lx == lxc && i >= csrow[lxc] && i < csrow[lxc + 1];
      if (!sorted_ok) atomicOr(P.err, NIXB200_ERR_UNSORTED);
      const T* ecell = erow + lxc * FC;

      T f6[6];
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const bool hz = (0x1C >> k) & 1, hy = (0x2A >> k) & 1, hx = (0x31 >> k) & 1;
        T     wz[N1], wy[N1], wx[N1];
#pragma unroll
        for (int j = 0; j < N1; j++) {
          wz[j] = hz ? wh[0][j] : wi[0][j];
          wy[j] = hy ? wh[1][j] : wi[1][j];
          wx[j] = hx ? wh[2][j] : wi[2][j];
        }
        const T* e = ecell + (hz ? bh[0] : 0) * esz + (hy ? bh[1] : 0) * esy + (hx ? bh[2] : 0) * FC + k;
        f6[k] = gather1<O, S, T>(e, wz, wy, wx);
      }
      T ex = mul<S>(f6[0], P.dt1), ey = mul<S>(f6[1], P.dt1), ez = mul<S>(f6[2], P.dt1);
      T bxx = mul<S>(f6[3], P.dt1), byy = mul<S>(f6[4], P.dt1), bzz = mul<S>(f6[5], P.dt1);

      // ---- momentum update: the reference's three pushers, association order preserved ---------
      T ux = mine[3 * 32], uy = mine[4 * 32], uz = mine[5 * 32];
      if (P.pusher == NIXB200_PUSH_BORIS) { // push_boris, primitives.hpp:165-189
        ux = add<S>(ux, ex);
        uy = add<S>(uy, ey);
        uz = add<S>(uz, ez);
        T gm = div_<S>(T(1.0), sqrt_<S>(add<S>(add<S>(add<S>(mul<S>(P.cc, P.cc), mul<S>(ux, ux)), mul<S>(uy, uy)), mul<S>(uz, uz))));
        bxx = mul<S>(bxx, gm);
        byy = mul<S>(byy, gm);
        bzz = mul<S>(bzz, gm);
        T bb = div_<S>(T(2.0), add<S>(add<S>(add<S>(T(1.0), mul<S>(bxx, bxx)), mul<S>(byy, byy)), mul<S>(bzz, bzz)));
        T vx = add<S>(ux, sub<S>(mul<S>(uy, bzz), mul<S>(uz, byy)));
        T vy = add<S>(uy, sub<S>(mul<S>(uz, bxx), mul<S>(ux, bzz)));
        T vz = add<S>(uz, sub<S>(mul<S>(ux, byy), mul<S>(uy, bxx)));
        ux = add<S>(ux, add<S>(mul<S>(sub<S>(mul<S>(vy, bzz), mul<S>(vz, byy)), bb), ex));
        uy = add<S>(uy, add<S>(mul<S>(sub<S>(mul<S>(vz, bxx), mul<S>(vx, bzz)), bb), ey));
        uz = add<S>(uz, add<S>(mul<S>(sub<S>(mul<S>(vx, byy), mul<S>(vy, bxx)), bb), ez));
      } else if (P.pusher == NIXB200_PUSH_VAY) { // push_vay, primitives.hpp:193-224
        T gm = div_<S>(T(1.0), sqrt_<S>(add<S>(add<S>(add<S>(mul<S>(P.cc, P.cc), mul<S>(ux, ux)), mul<S>(uy, uy)), mul<S>(uz, uz))));
        T vx = add<S>(add<S>(ux, mul<S>(T(2.0), ex)), mul<S>(gm, sub<S>(mul<S>(uy, bzz), mul<S>(uz, byy))));
        T vy = add<S>(add<S>(uy, mul<S>(T(2.0), ey)), mul<S>(gm, sub<S>(mul<S>(uz, bxx), mul<S>(ux, bzz))));
        T vz = add<S>(add<S>(uz, mul<S>(T(2.0), ez)), mul<S>(gm, sub<S>(mul<S>(ux, byy), mul<S>(uy, bxx))));
        gm        = add<S>(add<S>(add<S>(mul<S>(P.cc, P.cc), mul<S>(vx, vx)), mul<S>(vy, vy)), mul<S>(vz, vz));
        T bb = add<S>(add<S>(mul<S>(bxx, bxx), mul<S>(byy, byy)), mul<S>(bzz, bzz));
        T bu = add<S>(add<S>(mul<S>(bxx, vx), mul<S>(byy, vy)), mul<S>(bzz, vz));
        T xx = sub<S>(gm, bb);
        T yy = add<S>(bb, mul<S>(bu, bu));
        gm = div_<S>(T(1.0), sqrt_<S>(mul<S>(T(0.5), add<S>(xx, sqrt_<S>(add<S>(mul<S>(xx, xx), mul<S>(T(4.0), yy)))))));
        bxx = mul<S>(bxx, gm);
        byy = mul<S>(byy, gm);
        bzz = mul<S>(bzz, gm);
        bu  = add<S>(add<S>(mul<S>(bxx, vx), mul<S>(byy, vy)), mul<S>(bzz, vz));
        bb  = div_<S>(T(1.0), add<S>(add<S>(add<S>(T(1.0), mul<S>(bxx, bxx)), mul<S>(byy, byy)), mul<S>(bzz, bzz)));
        ux  = mul<S>(add<S>(add<S>(vx, mul<S>(bu, bxx)), sub<S>(mul<S>(vy, bzz), mul<S>(vz, byy))), bb);
        uy  = mul<S>(add<S>(add<S>(vy, mul<S>(bu, byy)), sub<S>(mul<S>(vz, bxx), mul<S>(vx, bzz))), bb);
        uz  = mul<S>(add<S>(add<S>(vz, mul<S>(bu, bzz)), sub<S>(mul<S>(vx, byy), mul<S>(vy, bxx))), bb);
      } else { // push_higuera_cary, primitives.hpp:227-253
        ux = add<S>(ux, ex);
        uy = add<S>(uy, ey);
        uz = add<S>(uz, ez);
        T gm = add<S>(add<S>(add<S>(mul<S>(P.cc, P.cc), mul<S>(ux, ux)), mul<S>(uy, uy)), mul<S>(uz, uz));
        T bb = add<S>(add<S>(mul<S>(bxx, bxx), mul<S>(byy, byy)), mul<S>(bzz, bzz));
        T bu = add<S>(add<S>(mul<S>(bxx, ux), mul<S>(byy, uy)), mul<S>(bzz, uz));
        T xx = sub<S>(gm, bb);
        T yy = add<S>(bb, mul<S>(bu, bu));
        gm = div_<S>(T(1.0), sqrt_<S>(mul<S>(T(0.5), add<S>(xx, sqrt_<S>(add<S>(mul<S>(xx, xx), mul<S>(T(4.0), yy)))))));
        bxx = mul<S>(bxx, gm);
        byy = mul<S>(byy, gm);
        bzz = mul<S>(bzz, gm);
        bb  = div_<S>(T(2.0), add<S>(add<S>(add<S>(T(1.0), mul<S>(bxx, bxx)), mul<S>(byy, byy)), mul<S>(bzz, bzz)));
        T vx = add<S>(ux, sub<S>(mul<S>(uy, bzz), mul<S>(uz, byy)));
        T vy = add<S>(uy, sub<S>(mul<S>(uz, bxx), mul<S>(ux, bzz)));
        T vz = add<S>(uz, sub<S>(mul<S>(ux, byy), mul<S>(uy, bxx)));
        ux = add<S>(ux, add<S>(mul<S>(sub<S>(mul<S>(vy, bzz), mul<S>(vz, byy)), bb), ex));
        uy = add<S>(uy, add<S>(mul<S>(sub<S>(mul<S>(vz, bxx), mul<S>(vx, bzz)), bb), ey));
        uz = add<S>(uz, add<S>(mul<S>(sub<S>(mul<S>(vx, byy), mul<S>(vy, bxx)), bb), ez));
      }

      // ---- position update (lorentz_factor, primitives.hpp:158-161); the old position goes to the
      //      temporary array like the reference's xv[0:3] = xu[0:3] (test_esirkepov.cpp:1046-1051)
      T uu  = add<S>(add<S>(mul<S>(ux, ux), mul<S>(uy, uy)), mul<S>(uz, uz));
      T gam = sqrt_<S>(add<S>(T(1.0), mul<S>(mul<S>(uu, P.rc), P.rc)));
      T dtg = div_<S>(P.delt, gam);
      T pn[3];
      pn[2] = add<S>(pos[2], mul<S>(ux, dtg));
      pn[1] = add<S>(pos[1], mul<S>(uy, dtg));
      pn[0] = add<S>(pos[0], mul<S>(uz, dtg));
      xv[soa(0, cap, i)] = pos[2];
      xv[soa(1, cap, i)] = pos[1];
      xv[soa(2, cap, i)] = pos[0];
      xu[soa(0, cap, i)] = pn[2];
      xu[soa(1, cap, i)] = pn[1];
      xu[soa(2, cap, i)] = pn[0];
      xu[soa(3, cap, i)] = ux;
      xu[soa(4, cap, i)] = uy;
      xu[soa(5, cap, i)] = uz;

      // ---- bin of the new position: count / classify ------------------------------------------
      int  i1[3];
      bool cfl_ok = true;
      int  dcode  = 0;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        i1[a]  = digitize(pn[a], c.off[a], P.rdel[a]);
        int dd = (pn[a] >= c.hi[a]) - (pn[a] < c.lo[a]) + 1;
        dcode  = dcode * 3 + dd;
        int sf = (i1[a] - g.is_odd) - ki[a];
        cfl_ok = cfl_ok && (sf >= -1) && (sf <= 1);
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
    }
    // ---- leavers: ordered rank inside (bin, direction); the counts of the bin that continues into
    //      the next iteration are carried in mydc ------------------------------------------------
    const bool     leaver = valid && dir != 13;
    const unsigned lm     = __ballot_sync(FULL, leaver);
    if (lm || carry) {
      const int last_lx = __shfl_sync(FULL, lxc, min(32, pe - i0) - 1);
      int       rk = 0, cnt = 0;
      if (leaver) {
        const unsigned grpm = __match_any_sync(lm, lxc * 32 + dir);
        rk                  = __popc(grpm & ((1u << lane) - 1));
        cnt                 = __popc(grpm);
        const int base      = (carry && lxc == open_lx) ? mydc[dir] : 0;
        const int rr        = base + rk;
        const int bb3[3]    = {bz, by, b0[2] + lxc};
        const int se        = slab_entry(g, dir, bb3);
        if (se < 0) atomicOr(P.err, NIXB200_ERR_CFL);
        int slot = atomicAdd(P.sp.nleave, 1);
        if (slot < P.sp.lcap && se >= 0) P.sp.lrec[slot] = make_int4(i, ch, se, (rr << 5) | dir);
        else atomicOr(P.err, NIXB200_ERR_CAPACITY);
        // leavers of (slab bin, direction) so far -> slab counts (scanned by k_mig_scan)
        if (rk == 0 && se >= 0) atomicMax(&P.sp.slabcnt[(size_t)ch * g.slaboff[27] + se], base + cnt);
      }
      __syncwarp();
      const bool keep = carry && last_lx == open_lx; // the open bin continues
      if (!keep && lane < 27) mydc[lane] = 0;
      open_lx = last_lx;
      __syncwarp();
      if (leaver && lxc == last_lx && rk == 0) mydc[dir] += cnt;
      carry = keep || __any_sync(FULL, leaver && lxc == last_lx);
      __syncwarp();
    }
  }
}


template <int O, bool S, typename T>
__global__ void __launch_bounds__(DTHREADS, (O >= 3) ? 2 : (sizeof(T) == 4 ? 4 : NIX_D_MINB)) k_deposit(const KparamsT<T> P)
{
  using C          = Cfg<O>;
  constexpr int N1 = C::N1;
  constexpr int NS = C::NS;
  constexpr int PV = C::PV;
  (void)NS;
  const Geo&    g    = P.geo;
  const int     tid  = threadIdx.x;
  const int     lane = tid & 31;
  const int     warp = tid >> 5;

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
  constexpr int JZ = C::JZ, JY = C::JY, JX = C::JX;
  constexpr int REC_D = (DWARPS * MAXMOV * CREC + 15) / 16 * 16;

  if (P.err[1]) return;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  T* smem_d = reinterpret_cast<T*>(smem_raw);
  T*   s_j    = smem_d;
  T*   s_rec  = s_j + C::J_DOUBLES;
  T*   s_red  = s_rec + REC_D;
  T*   s_pf   = s_red + DWARPS * C::SCR;
  int*      s_int  = reinterpret_cast<int*>(s_pf + DWARPS * PF_DOUBLES_W);
  int*      s_any  = s_int + ID_ANY;
  int*      s_next = s_int + ID_NEXT;
  int*      s_tbl  = s_int + ID_TBL;
  int*      s_mlst = s_int + ID_MLST;
  int*      s_cs   = s_int + ID_CS;
  ChunkGeoT<T>* s_cg = reinterpret_cast<ChunkGeoT<T>*>(s_int + ID_CG);

  const int32_t* __restrict__ start = P.sp.start;
  const int cellkey0 = ch * g.ncell;
  constexpr int CSW = C::TX + 1;
  if (tid == 0) {
    *s_any  = 0;
    *s_next = DWARPS;
  }
  stage_chunk_geo(s_cg, P.cg + ch, tid, DTHREADS);
  for (int t = tid; t < nbn[0] * nbn[1] * CSW; t += DTHREADS) {
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

  const int Lb  = g.nb;
  const int jz0 = b0[0] - g.is_odd - g.half + Lb - 1, jy0 = b0[1] - g.is_odd - g.half + Lb - 1,
            jx0 = b0[2] - g.is_odd - g.half + Lb - 1;

  for (int t = tid; t < C::J_DOUBLES; t += DTHREADS) s_j[t] = T(0.0);
  for (int t = tid; t < DWARPS * C::SCR; t += DTHREADS) s_red[t] = T(0.0);
  for (int v = tid; v < PV; v += DTHREADS) {
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

  const ChunkGeoT<T>& c = *s_cg;
  const size_t    cap = P.sp.cap;
  const T* __restrict__ xu = P.xu; // new positions
  const T* __restrict__ xv = P.xv; // old positions (written by k_push)
  T* myrec = s_rec + (size_t)warp * MAXMOV * CREC;
  MoverGeoT<T> mgeo;
#pragma unroll
  for (int a = 0; a < 3; a++) mgeo.del[a] = P.del[a], mgeo.rdel[a] = P.rdel[a];
  mgeo.is_odd = g.is_odd;
  int*    myml  = s_mlst + warp * MAXMOV;
  T* my_red = s_red + warp * C::SCR;
  T* my_pf  = s_pf + warp * PF_DOUBLES_W;
  const T* mine = my_pf + lane;

  // deposit role of the lane: particle slot ps, z-plane pl (lanes with pl >= N1 idle along)
  constexpr int PLW = C::PLW, NPS = C::NPS;
  const int     ps = lane / PLW, pl = lane % PLW;
  const bool    pl_on = pl < N1;
  const int     plc   = pl_on ? pl : N1 - 1;
  T acc[PV];
#pragma unroll
  for (int v = 0; v < PV; v++) acc[v] = T(0.0);
  // bin finished: sum the accumulators of the NPS slots through the (now idle) scratch -- lanes store
  // [value][lane], then lane g adds up the NPS entries of (value, plane) group g -- and add the sums
  // to the J tile.  Row stride 32 + N1: the strided 64-bit loads of 16 consecutive groups hit 16
  // different bank pairs.
  auto flush_bin = [&](int cellbase) {
    constexpr int RS  = 32 + N1;
    constexpr int NCH = (PV * RS + C::SCR - 1) / C::SCR; // chunks of values that fit the scratch
    constexpr int VCH = (PV + NCH - 1) / NCH;
    constexpr int G   = VCH * N1;
    static_assert(VCH * RS <= C::SCR, "value chunk must fit the scratch");
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
      __syncwarp();
#pragma unroll
      for (int vv = 0; vv < VCH; vv++)
        if (ch * VCH + vv < PV) my_red[vv * RS + lane] = acc[ch * VCH + vv];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < (G + 31) / 32; k++) {
        const int gi = lane + 32 * k;
        const int vv = gi / N1, plg = gi - vv * N1, v = ch * VCH + vv;
        if (gi < G && v < PV) {
          const T* src = my_red + vv * RS + plg;
          T        sum = T(0.0);
#pragma unroll
          for (int q = 0; q < NPS; q++) sum += src[q * PLW];
#if NIX_J_FIXED
          if (sum != T(0.0)) jt_add<T>(s_j + cellbase + (plg + 1) * JY * JX + s_tbl[v], sum, jt_scale<T>(P.q, P.qdxdt[0], P.qdxdt[1], P.qdxdt[2]));
#else
          if (sum != T(0.0)) atomicAdd(s_j + cellbase + (plg + 1) * JY * JX + s_tbl[v], sum);
#endif
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int v = 0; v < PV; v++) acc[v] = T(0.0);
  };

  const int ncell_t = nbn[0] * nbn[1] * nbn[2];
  auto advance = [&](int& cl, int& r, int& lx, int& i0, int& pe) -> bool {
    i0 += 32;
    while (i0 >= pe) {
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
  auto prefetch = [&](int i) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      cp_async_elem(my_pf + k * 32 + lane, xv + soa(k, cap, i));       // old x y z
      cp_async_elem(my_pf + (3 + k) * 32 + lane, xu + soa(k, cap, i)); // new x y z
    }
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
  if (have && ni0 + lane < npe) prefetch(ni0 + lane);
  cp_async_commit();

  int prev_cl = -1, nrec = 0;
  int bz = 0, by = 0, bx = 0, cellbase = 0;

  while (have) {
    const int cl = ncl, cr = nr, clx = nlx, i0 = ni0, pe = npe;
    have = advance(ncl, nr, nlx, ni0, npe);
    cp_async_wait_all();
    if (cl != prev_cl) { // first iteration of a bin
      const int lx = clx, lz = (nbn[1] == C::TY) ? cr / C::TY : cr / nbn[1], ly = cr - lz * nbn[1];
      bz = b0[0] + lz, by = b0[1] + ly, bx = b0[2] + lx;
      cellbase = (lz * JY + ly) * JX + lx; // J-tile node of mesh slot (0,0,0)
      prev_cl  = cl;
    }
    {
      const int  i     = i0 + lane;
      const bool valid = i < pe;
      bool       dep_ok = false, mover = false;
      int        mcode = -1; // single-axis mover: axis*2 + (1 if high side); -1: multi-axis
      T     pn[3] = {T(0.0), T(0.0), T(0.0)};
      T     wi[3][N1];
      int        ki[3] = {0, 0, 0}, sft[3] = {0, 0, 0};

      if (valid) {
        const T pos[3] = {mine[2 * 32], mine[1 * 32], mine[0]}; // old z,y,x
        pn[0] = mine[5 * 32], pn[1] = mine[4 * 32], pn[2] = mine[3 * 32];
        bool sorted_ok = true, cfl_ok = true;
        int  nmove = 0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const int ii = digitize(pos[a], c.off[a], P.rdel[a]);
          ki[a]        = ii - g.is_odd;
          shape_mc<O, S, T>(pos[a], add<S>(c.imin[a], mul<S>((T)ki[a], P.del[a])), P.rdel[a], wi[a]);
          sorted_ok = sorted_ok && (ii == ((a == 0) ? bz : ((a == 1) ? by : bx)));
          const int i1 = digitize(pn[a], c.off[a], P.rdel[a]);
          sft[a]       = (i1 - g.is_odd) - ki[a];
          cfl_ok       = cfl_ok && (sft[a] >= -1) && (sft[a] <= 1);
          if (sft[a] != 0) {
            nmove++;
            mcode = a * 2 + (sft[a] > 0 ? 1 : 0);
          }
        }
        dep_ok = cfl_ok && sorted_ok; // (both flags are raised by k_push)
        mover  = dep_ok && nmove > 0;
        if (nmove > 1) mcode = -1;
      }

      // ---- movers: compact record (old / new position, bin); the list is flushed when it is full ----
      unsigned mm = __ballot_sync(FULL, mover);
      while (mm) {
        const int  room = MAXMOV - nrec;
        const int  rk   = __popc(mm & ((1u << lane) - 1));
        const bool take = mover && ((mm >> lane) & 1u) && rk < room;
        if (take) {
          T* r = myrec + (nrec + rk) * CREC;
          r[0] = mine[2 * 32], r[1] = mine[1 * 32], r[2] = mine[0]; // old z, y, x
          r[3] = pn[0], r[4] = pn[1], r[5] = pn[2];                   // new z, y, x
          int* ri = reinterpret_cast<int*>(r + 6);
          ri[0]   = cellbase;
          ri[1]   = mcode;
          ri[2]   = bz;
          ri[3]   = by | (bx << 16);
        }
        const unsigned taken = __ballot_sync(FULL, take);
        nrec += __popc(taken);
        mm &= ~taken;
        if (mm || nrec >= NIX_MOV_FLUSH) { // full batches: whole groups of XGROUP records keep the flush lanes busy
          flush_movers<O, S, T>(s_j, myrec, my_red, myml, nrec, s_cg, mgeo, P.q, P.qdxdt[0], P.qdxdt[1], P.qdxdt[2]);
          nrec = 0;
        }
      }
      __syncwarp();
      // the staged positions have been consumed: request the next iteration's
      if (have && ni0 + lane < npe) prefetch(ni0 + lane);
      cp_async_commit();

      // ---- 1-D deposit weights, axis by axis straight into the warp's scratch --------------------
      // ss[0][.][1..O+1] = old weights; ss[1][.][1+sft..] = new weights (test_esirkepov.cpp:1060-1085);
      // a particle that may not deposit (out of its bin / more than one bin per step) leaves zeros
      {
        using R2 = typename Real<T>::vec2;
        R2* q2 = reinterpret_cast<R2*>(my_red);
#pragma unroll
        for (int a = 0; a < 3; a++) {
          T s0[N1], ds[N1], cp[N1];
          if (valid && dep_ok) {
            T wn[N1];
            int    k1 = ki[a] + sft[a];
            shape_mc<O, S, T>(pn[a], add<S>(c.imin[a], mul<S>((T)k1, P.del[a])), P.rdel[a], wn);
            T run = (sft[a] < 0) ? wn[0] : T(0.0); // DS of mesh slot 0 (ds3d, esirkepov.hpp:167-174)
#pragma unroll
            for (int j = 0; j < N1; j++) { // central slots j + 1
              const T vm = (j + 1 <= O) ? wn[j + 1] : T(0.0); // sft = -1 : slot <- wn[slot]
              const T v0 = wn[j];                           // sft =  0
              const T vp = (j >= 1) ? wn[j - 1] : T(0.0);      // sft = +1
              const T s1 = (sft[a] == 0) ? v0 : ((sft[a] < 0) ? vm : vp);
              s0[j] = wi[a][j];
              ds[j] = s1 - s0[j];
              cp[j] = run;
              run += ds[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < N1; j++) s0[j] = ds[j] = cp[j] = T(0.0);
          }
          if (a == 0) {
#pragma unroll
            for (int z = 0; z < N1; z++) {
              q2[C::WQ_Z / 2 + z * C::ZP + lane] = Real<T>::make2(s0[z], ds[z]);
              if (z >= 1) my_red[C::WQ_ZC + (z - 1) * C::ZP + lane] = cp[z];
            }
          } else {
            constexpr int NPR = C::NPR;
            T        t[2 * NPR];
#pragma unroll
            for (int j = 0; j < N1; j++) {
              t[j]      = s0[j];
              t[N1 + j] = ds[j];
              if (j >= 1) t[2 * N1 + j - 1] = cp[j];
            }
            if (3 * N1 - 1 < 2 * NPR) t[2 * NPR - 1] = T(0.0);
#pragma unroll
            for (int pr = 0; pr < NPR; pr++) q2[((a - 1) * NPR + pr) * 32 + lane] = Real<T>::make2(t[2 * pr], t[2 * pr + 1]);
          }
        }
      }
      __syncwarp();
      {
        const int nround = (min(32, pe - i0) + NPS - 1) / NPS;
        for (int r = 0; r < nround; r++) deposit_round<O, T>(my_red, r * NPS + ps, plc, pl_on, P.q, P.qdxdt, acc);
      }
      __syncwarp();
    }
    if (!have || ncl != cl) flush_bin(cellbase); // last iteration of a bin: its current -> J tile
  }
  if (nrec) flush_movers<O, S, T>(s_j, myrec, my_red, myml, nrec, s_cg, mgeo, P.q, P.qdxdt[0], P.qdxdt[1], P.qdxdt[2]);

  // ---- flush the J tile: the CTA's single scatter to global memory --------------------------------
  __syncthreads();
  T* __restrict__ ujc = P.uj + (size_t)ch * g.M[0] * g.M[1] * g.M[2] * 4;
  for (int t = tid; t < JZ * JY * JX * 4; t += DTHREADS) {
#if NIX_J_FIXED
    const T v = jt_read<T>(s_j + (t & 3) * C::JC + (t >> 2), T(1.0) / jt_scale<T>(P.q, P.qdxdt[0], P.qdxdt[1], P.qdxdt[2]));
#else
    const T v = s_j[(t & 3) * C::JC + (t >> 2)];
#endif
    if (v != T(0.0)) {
      const int k = t & 3, n = t >> 2;
      const int gx = jx0 + n % JX, gy = jy0 + (n / JX) % JY, gz = jz0 + n / (JX * JY);
      if (gx >= 0 && gx < g.M[2] && gy >= 0 && gy < g.M[1] && gz >= 0 && gz < g.M[0])
        atomicAdd(&ujc[(((size_t)gz * g.M[1] + gy) * g.M[2] + gx) * 4 + k], v);
    }
  }
}

template <int O, bool S, typename T>
int launch_split_t(const PushArgs& a, const CUtensorMap* tmap, cudaStream_t st, cudaEvent_t* ev)
{
  KparamsT<T> P;
  P.geo  = a.geo;
  P.cg   = a.cg;
  P.uj   = reinterpret_cast<T*>(a.uj);
  P.xu   = reinterpret_cast<T*>(a.sp.xu);
  P.xv   = reinterpret_cast<T*>(a.sp.xv);
  P.sp   = a.sp;
  P.delt = (T)a.delt;
  P.dt1  = (T)(0.5 * a.sp.q / a.sp.m * a.delt); // ref_driver.cpp: dt1
  P.q    = (T)a.sp.q;
  for (int d = 0; d < 3; d++) {
    P.qdxdt[d] = (T)(a.sp.q * (a.geo.del[d] / a.delt));
    P.del[d]   = (T)a.geo.del[d];
    P.rdel[d]  = (T)a.geo.rdel[d];
  }
  P.cc  = (T)a.geo.cc;
  P.rc  = (T)a.geo.rc;
  P.err = a.err;
  P.pusher = a.pusher;
  int nblocks = a.geo.nchunk * a.geo.ntile;
  if (ev) cudaEventRecord(ev[0], st);
  k_push<O, S, T><<<nblocks, PTHREADS, push_smem<O, T>(), st>>>(*tmap, P);
  NIX_LAUNCHED();
  if (ev) {
    cudaEventRecord(ev[1], st);
    cudaEventRecord(ev[2], st);
  }
  k_deposit<O, S, T><<<nblocks, DTHREADS, deposit_smem<O, T>(), st>>>(P);
  NIX_LAUNCHED();
  if (ev) cudaEventRecord(ev[3], st);
  return 0;
}

template <int O, bool S, typename T>
int prepare_t()
{
  NIX_CUDA(cudaFuncSetAttribute(k_push<O, S, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  NIX_CUDA(cudaFuncSetAttribute(k_deposit<O, S, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  return 0;
}
} // namespace

// The dynamic shared-memory limit is a per-DEVICE attribute of a kernel: set for the current device by
// every nixb200_domain_create (cheap; several domains on several devices may live in one process).
int push_deposit_prepare(int order, bool fp32)
{
  if (fp32) {
    switch (order) {
    case 1: return prepare_t<1, false, float>();
    case 2: return prepare_t<2, false, float>();
    case 3: return prepare_t<3, false, float>();
    default: set_error("order must be 1, 2 or 3"); return 1;
    }
  }
  switch (order) {
  case 1: return prepare_t<1, true, double>() || prepare_t<1, false, double>();
  case 2: return prepare_t<2, true, double>() || prepare_t<2, false, double>();
  case 3: return prepare_t<3, true, double>() || prepare_t<3, false, double>();
  default: set_error("order must be 1, 2 or 3"); return 1;
  }
}

size_t push_smem_bytes(const Geo& g)
{
  switch (g.order) {
  case 1: return std::max(push_smem<1, double>(), deposit_smem<1, double>());
  case 2: return std::max(push_smem<2, double>(), deposit_smem<2, double>());
  default: return std::max(push_smem<3, double>(), deposit_smem<3, double>());
  }
}

// box of the E/B tile the push kernel stages (nodes per axis z, y, x; dummy rows / columns included)
void push_tile_box(int order, int& ez, int& ey, int& ex)
{
  switch (order) {
  case 1: ez = Cfg<1>::EZ, ey = Cfg<1>::EY, ex = Cfg<1>::EX; break;
  case 2: ez = Cfg<2>::EZ, ey = Cfg<2>::EY, ex = Cfg<2>::EX; break;
  default: ez = Cfg<3>::EZ, ey = Cfg<3>::EY, ex = Cfg<3>::EX; break;
  }
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

int launch_push_deposit(const PushArgs& a, const CUtensorMap* tmap, bool strict, cudaStream_t st, cudaEvent_t* ev)
{
  // leaver bookkeeping of this step
  NIX_CUDA(cudaMemsetAsync(a.sp.slabcnt, 0, sizeof(int32_t) * (size_t)a.geo.nchunk * a.geo.slaboff[27], st));
  NIX_CUDA(cudaMemsetAsync(a.sp.oob, 0, sizeof(int32_t) * a.geo.nchunk * LANES, st));
  NIX_CUDA(cudaMemsetAsync(a.sp.nleave, 0, sizeof(int32_t), st));
  if (a.fp32) { // the fp32 mode has no bit-exact reference to follow: contracted arithmetic only
    switch (a.geo.order) {
    case 1: return launch_split_t<1, false, float>(a, tmap, st, ev);
    case 2: return launch_split_t<2, false, float>(a, tmap, st, ev);
    case 3: return launch_split_t<3, false, float>(a, tmap, st, ev);
    default: set_error("order must be 1, 2 or 3"); return 1;
    }
  }
  switch (a.geo.order) {
  case 1: return strict ? launch_split_t<1, true, double>(a, tmap, st, ev) : launch_split_t<1, false, double>(a, tmap, st, ev);
  case 2: return strict ? launch_split_t<2, true, double>(a, tmap, st, ev) : launch_split_t<2, false, double>(a, tmap, st, ev);
  case 3: return strict ? launch_split_t<3, true, double>(a, tmap, st, ev) : launch_split_t<3, false, double>(a, tmap, st, ev);
  default: set_error("order must be 1, 2 or 3"); return 1;
  }
}
} // namespace nixb200
