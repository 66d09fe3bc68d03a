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
// Work item = one x-segment of one (z,y) row of cells of one chunk; its particles are contiguous
// because the container is cell-sorted.  One CTA of 128 threads per item:
//   * TMA (cp.async.bulk.tensor.5d) stages the (O+2)x(O+2)x(SEG+O+1)x6 E/B tile of the row -- ghosts
//     included -- into shared memory; an mbarrier signals arrival.
//   * phase 1, thread per particle (batches of 128): coalesced SoA loads, weights, gather from the
//     smem tile over the EXACT support of every component ((O+1)^3 points; the half-grid components
//     start one node later when the particle sits in the upper half of its cell), Boris, move,
//     coalesced stores, bin of the new position (-> key + histogram, or a leaver record with its
//     ordered rank inside the item), and the 1-D deposit weights of the particle into a
//     particle-major scratch row (128-bit stores, conflict free).
//   * phase 2 accumulates the Esirkepov current of all particles of one cell in REGISTERS (the GPU
//     analogue of the reference's sorted `reduce_add` path, primitives.hpp:798-809), thread per
//     (jy,jz) column of the deposit mesh, particles of the cell split over thread groups:
//       F path  particles that stay in their cell (the vast majority): their weights touch only the
//               central (O+1)^3 nodes, so a group is (O+1)^2 threads and the x loop has O+1 slots;
//               the accumulators slide along x with the cell index and live across batches.
//       C path  particles that change cell: full (O+3)^3 mesh, (O+3)^2 threads per group, transient
//               accumulators flushed into a small sliding window in shared memory.
//     When a cell is finished the x-slot that leaves the window is reduced over the groups through
//     a staging buffer, merged with the C window and added to J in global memory with
//     red.global.add.f64 -- once per (z,y,x,component) and CTA, no shared-memory atomics.
#include "common.cuh"

namespace nixb200
{
namespace
{
constexpr int THREADS = 128;
constexpr int NWARP   = THREADS / 32;
constexpr int MAXSEG  = 48;

template <int O>
struct Cfg {
  static constexpr int NW   = O + 2;          // gather stencil width of the staged tile (interp.hpp)
  static constexpr int N1   = O + 1;          // support of one shape function
  static constexpr int NS   = O + 3;          // deposit mesh width     (esirkepov.hpp:326-328)
  static constexpr int NCOL = NS * NS;        // (jy,jz) columns of the full mesh       (C path)
  static constexpr int GC   = THREADS / NCOL; // C groups: 8 / 5 / 3
  static constexpr int NCF  = N1 * N1;        // (jy,jz) columns of the central mesh    (F path)
  static constexpr int GF   = THREADS / NCF;  // F groups: 32 / 14 / 8
  // particle-major scratch row (doubles); pairs are 16-byte aligned
  static constexpr int PY  = 0;               // (S0y[j], DSy[j]) j = 1..N1
  static constexpr int PZ  = 2 * N1;
  static constexpr int PX  = 4 * N1;
  static constexpr int CY  = 6 * N1;          // CPy[j] = sum of DSy over slots < j, j = 1..N1
  static constexpr int CZ  = 7 * N1;
  static constexpr int OUT = 8 * N1;          // DSy0 DSyL CPyL DSz0 DSzL CPzL DSx0 DSxL  (L = NS-1)
  static constexpr int ROW = 8 * N1 + 10;     // == 2 (mod 4): 128-bit stores of 8 lanes hit 32 banks
};

struct SmemLayout {
  int    eb_doubles, stf_doubles, win_doubles, scratch_doubles;
  size_t bytes;
};

constexpr int SMEM_INTS = 1024;

template <int O>
__host__ __device__ inline SmemLayout smem_layout(int seg)
{
  using C = Cfg<O>;
  SmemLayout L;
  int        ex     = seg + C::NW - 1;
  L.eb_doubles      = ((C::NW * C::NW * ex * 6) + 15) / 16 * 16; // keep 128-byte multiples
  L.stf_doubles     = ((2 * C::GF * C::NCF * 4) + 15) / 16 * 16; // double buffered
  L.win_doubles     = ((C::NS * C::NCOL * 4) + 15) / 16 * 16;
  L.scratch_doubles = (C::ROW * THREADS + 15) / 16 * 16;
  L.bytes = sizeof(double) * ((size_t)L.eb_doubles + L.stf_doubles + L.win_doubles + L.scratch_doubles) +
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
// e points at (bz, by, tx + bx, component); EX = x extent of the staged tile.
template <int O, bool S>
__device__ __forceinline__ double gather1(const double* __restrict__ e, int EX, const double* wz,
                                          const double* wy, const double* wx)
{
  constexpr int NW = O + 2;
  double        rz = 0.0;
#pragma unroll
  for (int jz = 0; jz <= O; jz++) {
    double ry = 0.0;
#pragma unroll
    for (int jy = 0; jy <= O; jy++) {
      double        rx = 0.0;
      const double* p  = e + (size_t)((jz * NW + jy) * EX) * 6;
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

template <int O, bool S>
__global__ void __launch_bounds__(THREADS, 3) k_push_deposit(const __grid_constant__ CUtensorMap tmap,
                                                             const Kparams P)
{
  using C            = Cfg<O>;
  constexpr int NW   = C::NW;
  constexpr int N1   = C::N1;
  constexpr int NS   = C::NS;
  constexpr int NCOL = C::NCOL;
  constexpr int GC   = C::GC;
  constexpr int NCF  = C::NCF;
  constexpr int GF   = C::GF;
  constexpr int ROW  = C::ROW;
  const Geo&    g    = P.geo;
  const int     tid  = threadIdx.x;
  const int     lane = tid & 31;
  const int     warp = tid >> 5;

  // ---- decode the work item -------------------------------------------------------------------
  const int item = blockIdx.x % g.nitem;
  const int ch   = blockIdx.x / g.nitem;
  const int sg   = item % g.nseg;
  const int ry   = (item / g.nseg) % g.R[1];
  const int rz   = item / (g.nseg * g.R[1]);
  const int xs   = sg * g.seg;              // first cell (bin index) of the segment
  const int ncs  = min(g.seg, g.R[2] - xs); // cells in this segment
  const int row0 = (ch * g.ncell + (rz * g.R[1] + ry) * g.R[2] + xs) * LANES; // key of first bin

  const int32_t* __restrict__ start = P.sp.start;
  const int p_begin = start[row0];
  const int p_end   = start[row0 + ncs * LANES];
  if (p_begin == p_end) return;

  // ---- shared memory carve-up (indices into one __shared__ array: keeps LDS/STS addressing) -----
  extern __shared__ __align__(1024) double smem_d[];
  const SmemLayout L = smem_layout<O>(g.seg);
  double*   s_eb      = smem_d;
  double*   s_stf     = smem_d + L.eb_doubles;
  double*   s_win     = s_stf + L.stf_doubles;
  double*   s_scr     = s_win + L.win_doubles;
  int*      s_int     = reinterpret_cast<int*>(s_scr + L.scratch_doubles);
  uint64_t* s_bar     = reinterpret_cast<uint64_t*>(s_int); // 2 ints
  int*      s_pidx    = s_int + 2;                          // [seg+1]
  int*      s_dirbase = s_int + 64;                         // [27]
  int*      s_warpcnt = s_int + 96;                         // [4][27]
  int*      s_ccnt    = s_int + 208;                        // [2][MAXSEG] movers per cell, by batch parity
  int*      s_wcnt    = s_int + 304;                        // [4] movers per warp
  int*      s_cls     = s_int + 320;                        // [128] 1 = mover (C path)
  int*      s_clist   = s_int + 448;                        // [4][32] batch slots of the movers, per warp

  const int EX = g.seg + NW - 1; // E/B tile extent along x

  // tile origins (array indices)
  const int Lb  = g.nb;
  const int ez0 = rz - g.is_odd - g.half + Lb, ey0 = ry - g.is_odd - g.half + Lb,
            ex0 = xs - g.is_odd - g.half + Lb;
  const int jz0 = ez0 - 1, jy0 = ey0 - 1, jx0 = ex0 - 1;

  if (tid == 0) {
    mbar_init(s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int t = tid; t <= ncs; t += THREADS) s_pidx[t] = start[row0 + t * LANES];
  for (int t = tid; t < 27; t += THREADS) s_dirbase[t] = 0;
  for (int t = tid; t < 4 * 27; t += THREADS) s_warpcnt[t] = 0;
  for (int t = tid; t < 2 * MAXSEG; t += THREADS) s_ccnt[t] = 0;
  for (int t = tid; t < NS * NCOL * 4; t += THREADS) s_win[t] = 0.0;
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(s_bar, (uint32_t)(NW * NW * EX * 6 * sizeof(double)));
    tma_load_5d(s_eb, &tmap, s_bar, 0, ex0, ey0, ez0, ch);
  }

  const ChunkGeo& c   = P.cg[ch];
  const int       cb  = P.sp.cbase[ch];
  const size_t    cap = P.sp.cap;
  double* __restrict__ xu = P.sp.xu;
  double* __restrict__ ujc = P.uj + (size_t)ch * g.M[0] * g.M[1] * g.M[2] * 4;

  // ---- phase-2 roles ---------------------------------------------------------------------------
  // F: group gf, central column (fy, fz) in 0..O  <->  mesh column (fy+1, fz+1)
  const bool f_active = tid < GF * NCF;
  const int  gf       = tid / NCF;
  const int  colf     = tid % NCF;
  const int  fy       = colf % N1;
  const int  fz       = colf / N1;
  // C: group gc, mesh column (jy, jz) in 0..NS-1
  const bool c_active = tid < GC * NCOL;
  const int  gc       = tid / NCOL;
  const int  col      = tid % NCOL;
  const int  jy       = col % NS;
  const int  jz       = col / NS;
  // scratch offsets of this C column: (S0, DS) pair or lone DS, and CP; -1 = identically zero
  const int c_py = (jy >= 1 && jy <= N1) ? C::PY + 2 * (jy - 1) : -1;
  const int c_dy = (jy == 0) ? C::OUT + 0 : C::OUT + 1;
  const int c_cy = (jy == 0) ? -1 : ((jy <= N1) ? C::CY + jy - 1 : C::OUT + 2);
  const int c_pz = (jz >= 1 && jz <= N1) ? C::PZ + 2 * (jz - 1) : -1;
  const int c_dz = (jz == 0) ? C::OUT + 3 : C::OUT + 4;
  const int c_cz = (jz == 0) ? -1 : ((jz <= N1) ? C::CZ + jz - 1 : C::OUT + 5);

  // F accumulators: slot 0 carries the previous cell's slot 1, slots 1..N1 receive this cell
  double accf[N1 + 1][4];
#pragma unroll
  for (int s = 0; s <= N1; s++)
#pragma unroll
    for (int k = 0; k < 4; k++) accf[s][k] = 0.0;
  int cc   = 0; // current cell (relative to xs) of the sliding deposit window
  int par  = 0; // staging buffer parity
  int rot  = 0; // C window rotation: mesh slot s lives at plane (s + rot) % NS
  int bpar = 0; // batch parity of s_ccnt

  mbar_wait(s_bar, 0);

  // Retire slot 0 of the window: reduce the F accumulators over the GF groups through the staging
  // buffer, merge the C window, add to J in global memory (every (z,y,x) of the CTA's footprint is
  // retired exactly once, so this is the CTA's single flush of that entry), then slide by one cell.
  auto retire = [&]() {
    double* st = s_stf + par * (GF * NCF * 4);
    if (f_active) {
      double2* o = reinterpret_cast<double2*>(st + (gf * NCF + colf) * 4);
      o[0]       = make_double2(accf[0][0], accf[0][1]);
      o[1]       = make_double2(accf[0][2], accf[0][3]);
    }
    __syncthreads();
    double* w0 = s_win + (rot % NS) * (NCOL * 4);
    for (int t = tid; t < NCOL * 4; t += THREADS) {
      const int cl = t >> 2, k = t & 3;
      const int yy = cl % NS, zz = cl / NS;
      double    sum = w0[t];
      w0[t]         = 0.0;
      if (yy >= 1 && yy <= N1 && zz >= 1 && zz <= N1) {
        const double* sp = st + ((zz - 1) * N1 + (yy - 1)) * 4 + k;
#pragma unroll 4
        for (int gg = 0; gg < GF; gg++) sum += sp[gg * NCF * 4];
      }
      if (sum != 0.0) {
        int gy = jy0 + yy, gz = jz0 + zz, gx = jx0 + cc;
        if (gx >= 0 && gx < g.M[2] && gy >= 0 && gy < g.M[1] && gz >= 0 && gz < g.M[0])
          atomicAdd(&ujc[(((size_t)gz * g.M[1] + gy) * g.M[2] + gx) * 4 + k], sum);
      }
    }
    par ^= 1; // the other buffer is free: its readers passed the barrier above
    rot = (rot + 1) % NS;
#pragma unroll
    for (int s = 0; s < N1; s++)
#pragma unroll
      for (int k = 0; k < 4; k++) accf[s][k] = accf[s + 1][k];
#pragma unroll
    for (int k = 0; k < 4; k++) accf[N1][k] = 0.0;
    cc++;
  };

  for (int b0 = p_begin; b0 < p_end; b0 += THREADS) {
    const int  b1    = min(b0 + THREADS, p_end);
    const int  i     = b0 + tid;
    const bool valid = i < b1;
    int        dir   = 13;
    bool       mover = false;
    int*       ccnt  = s_ccnt + bpar * MAXSEG;

    // =============================== phase 1: push ===============================
    if (valid) {
      double pos[3], u[3]; // index 0,1,2 = z,y,x
      pos[2] = xu[soa(0, cap, i)];
      pos[1] = xu[soa(1, cap, i)];
      pos[0] = xu[soa(2, cap, i)];
      u[2]   = xu[soa(3, cap, i)];
      u[1]   = xu[soa(4, cap, i)];
      u[0]   = xu[soa(5, cap, i)];

      int    ki[3], sh[3], bh[3];
      double wi[3][N1], wh[3][N1];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        int ii = digitize(pos[a], c.off[a], g.rdel[a]);
        ki[a]  = ii - g.is_odd;
        int hh = digitize(pos[a], c.hoff[a], g.rdel[a]);
        shape_mc<O, S>(pos[a], add<S>(c.imin[a], mul<S>((double)ki[a], g.del[a])), g.rdel[a], wi[a]);
        shape_mc<O, S>(pos[a], add<S>(c.lo[a], mul<S>((double)hh, g.del[a])), g.rdel[a], wh[a]);
        // interp::shift_weights<O>(hh - ki, wh)  interp.hpp:154-160: the half-grid support starts
        // one node later; kept as a base offset instead of moving the weights
        bh[a] = (hh - ki[a] > 0) ? 1 : 0;
        sh[a] = ii;
      }
      const int  txo       = sh[2] - xs; // cell offset inside the segment
      const bool sorted_ok = (sh[0] == rz) && (sh[1] == ry) && (txo >= 0) && (txo < ncs);
      if (!sorted_ok) atomicOr(P.err, NIXB200_ERR_UNSORTED);
      const int tx = sorted_ok ? txo : 0;

      // ---- gather: Ex Ey Ez Bx By Bz; half-grid axes: Ex x | Ey y | Ez z | Bx y,z | By x,z | Bz x,y
      const double* e0 = s_eb + (size_t)tx * 6;
      auto          at = [&](int bz, int by, int bx, int k) { return e0 + (size_t)((bz * NW + by) * EX + bx) * 6 + k; };
      double rz6[6];
      rz6[0] = gather1<O, S>(at(0, 0, bh[2], 0), EX, wi[0], wi[1], wh[2]);
      rz6[1] = gather1<O, S>(at(0, bh[1], 0, 1), EX, wi[0], wh[1], wi[2]);
      rz6[2] = gather1<O, S>(at(bh[0], 0, 0, 2), EX, wh[0], wi[1], wi[2]);
      rz6[3] = gather1<O, S>(at(bh[0], bh[1], 0, 3), EX, wh[0], wh[1], wi[2]);
      rz6[4] = gather1<O, S>(at(bh[0], 0, bh[2], 4), EX, wh[0], wi[1], wh[2]);
      rz6[5] = gather1<O, S>(at(0, bh[1], bh[2], 5), EX, wi[0], wh[1], wh[2]);
      double ex = mul<S>(rz6[0], P.dt1), ey = mul<S>(rz6[1], P.dt1), ez = mul<S>(rz6[2], P.dt1);
      double bx = mul<S>(rz6[3], P.dt1), by = mul<S>(rz6[4], P.dt1), bz = mul<S>(rz6[5], P.dt1);

      // ---- push_boris (primitives.hpp:165-189) ----------------------------------------------
      double ux = u[2], uy = u[1], uz = u[0];
      ux = add<S>(ux, ex);
      uy = add<S>(uy, ey);
      uz = add<S>(uz, ez);
      double gm = div_<S>(1.0, sqrt_<S>(add<S>(add<S>(add<S>(mul<S>(g.cc, g.cc), mul<S>(ux, ux)), mul<S>(uy, uy)), mul<S>(uz, uz))));
      bx = mul<S>(bx, gm);
      by = mul<S>(by, gm);
      bz = mul<S>(bz, gm);
      double bb = div_<S>(2.0, add<S>(add<S>(add<S>(1.0, mul<S>(bx, bx)), mul<S>(by, by)), mul<S>(bz, bz)));
      double vx = add<S>(ux, sub<S>(mul<S>(uy, bz), mul<S>(uz, by)));
      double vy = add<S>(uy, sub<S>(mul<S>(uz, bx), mul<S>(ux, bz)));
      double vz = add<S>(uz, sub<S>(mul<S>(ux, by), mul<S>(uy, bx)));
      ux = add<S>(ux, add<S>(mul<S>(sub<S>(mul<S>(vy, bz), mul<S>(vz, by)), bb), ex));
      uy = add<S>(uy, add<S>(mul<S>(sub<S>(mul<S>(vz, bx), mul<S>(vx, bz)), bb), ey));
      uz = add<S>(uz, add<S>(mul<S>(sub<S>(mul<S>(vx, by), mul<S>(vy, bx)), bb), ez));

      // ---- position update (lorentz_factor, primitives.hpp:158-161) ------------------------
      double uu  = add<S>(add<S>(mul<S>(ux, ux), mul<S>(uy, uy)), mul<S>(uz, uz));
      double gam = sqrt_<S>(add<S>(1.0, mul<S>(mul<S>(uu, g.rc), g.rc)));
      double dtg = div_<S>(P.delt, gam);
      double pn[3];
      pn[2] = add<S>(pos[2], mul<S>(ux, dtg));
      pn[1] = add<S>(pos[1], mul<S>(uy, dtg));
      pn[0] = add<S>(pos[0], mul<S>(uz, dtg));

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
      bool moved  = false;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        i1[a]   = digitize(pn[a], c.off[a], g.rdel[a]);
        int dd  = (pn[a] >= c.hi[a]) - (pn[a] < c.lo[a]) + 1;
        dcode   = dcode * 3 + dd;
        int sft = (i1[a] - g.is_odd) - ki[a];
        cfl_ok  = cfl_ok && (sft >= -1) && (sft <= 1);
        moved   = moved || (sft != 0);
      }
      dir            = dcode;
      const int lnid = (i - cb) & (LANES - 1);
      if (dir == 13) {
        int key     = (ch * g.ncell + (i1[0] * g.R[1] + i1[1]) * g.R[2] + i1[2]) * LANES + lnid;
        P.sp.key[i] = key;
        atomicAdd(&P.sp.hist[key], 1);
      } else {
        P.sp.key[i] = -1;
        atomicAdd(&P.sp.oob[ch * LANES + lnid], 1);
      }
      if (!cfl_ok) atomicOr(P.err, NIXB200_ERR_CFL);

      // ---- 1-D deposit weights of this particle -> scratch row --------------------------------
      // ss[0][.][1..O+1] = old weights; ss[1][.][1+sft..] = new weights (test_esirkepov.cpp:1060-1085)
      // ds = ss[1] - ss[0] (ds3d, esirkepov.hpp:167-174); cp[j] = sum of ds over slots < j
      const bool dep_ok = cfl_ok && sorted_ok;
      mover             = dep_ok && moved;
      double* row       = s_scr + (size_t)tid * ROW;
      double  tail[2 * N1 + 8]; // CPy[1..N1] CPz[1..N1] | DSy0 DSyL CPyL DSz0 DSzL CPzL DSx0 DSxL
#pragma unroll
      for (int a = 0; a < 3; a++) {
        double wn[N1];
        int    k1  = i1[a] - g.is_odd;
        int    sft = k1 - ki[a];
        shape_mc<O, S>(pn[a], add<S>(c.imin[a], mul<S>((double)k1, g.del[a])), g.rdel[a], wn);
        const int base_p = (a == 2) ? C::PX : (a == 1 ? C::PY : C::PZ);
        double    cp     = 0.0;
        double    cpv[NS], dsv[NS], s0v[NS];
#pragma unroll
        for (int j = 0; j < NS; j++) {
          double s0 = (j >= 1 && j <= O + 1) ? wi[a][j - 1] : 0.0;
          double vm = (j >= 0 && j <= O) ? wn[j] : 0.0;         // sft = -1 : slot j <- wn[j]
          double v0 = (j >= 1 && j <= O + 1) ? wn[j - 1] : 0.0; // sft =  0
          double vp = (j >= 2 && j <= O + 2) ? wn[j - 2] : 0.0; // sft = +1
          double s1 = (sft == 0) ? v0 : ((sft < 0) ? vm : vp);
          if (!dep_ok) {
            s0 = 0.0;
            s1 = 0.0;
          }
          s0v[j] = s0;
          dsv[j] = s1 - s0;
          cpv[j] = cp;
          cp += dsv[j];
        }
#pragma unroll
        for (int j = 1; j <= N1; j++)
          *reinterpret_cast<double2*>(row + base_p + 2 * (j - 1)) = make_double2(s0v[j], dsv[j]);
        if (a != 2) {
          const int tc = (a == 1) ? 0 : N1;
          const int to = 2 * N1 + ((a == 1) ? 0 : 3);
#pragma unroll
          for (int j = 1; j <= N1; j++) tail[tc + j - 1] = cpv[j];
          tail[to + 0] = dsv[0];
          tail[to + 1] = dsv[NS - 1];
          tail[to + 2] = cpv[NS - 1];
        } else {
          tail[2 * N1 + 6] = dsv[0];
          tail[2 * N1 + 7] = dsv[NS - 1];
        }
      }
#pragma unroll
      for (int j = 0; j < N1 + 4; j++)
        *reinterpret_cast<double2*>(row + C::CY + 2 * j) = make_double2(tail[2 * j], tail[2 * j + 1]);
      s_cls[tid] = mover ? 1 : 0;
      if (mover) atomicAdd(&ccnt[tx], 1);
    }

    // ---- movers of this batch, in particle order (per-warp lists) ------------------------------
    {
      const unsigned mm = __ballot_sync(0xffffffffu, mover);
      if (mover) s_clist[warp * 32 + __popc(mm & ((1u << lane) - 1))] = tid;
      if (lane == 0) s_wcnt[warp] = __popc(mm);
    }

    // ---- ordered rank of the leavers inside this work item (segmented counting scan) ----------
    const bool leaver = valid && dir != 13;
    const int  any    = __syncthreads_or(leaver ? 1 : 0); // also: scratch of the batch is complete
    if (any) {
      unsigned lm     = __ballot_sync(0xffffffffu, leaver);
      int      rank_w = 0;
      if (leaver) {
        unsigned grpm = __match_any_sync(lm, dir);
        rank_w        = __popc(grpm & ((1u << lane) - 1));
        if (rank_w == 0) s_warpcnt[warp * 27 + dir] = __popc(grpm);
      }
      __syncthreads();
      if (leaver) {
        int r = s_dirbase[dir] + rank_w;
        for (int w = 0; w < warp; w++) r += s_warpcnt[w * 27 + dir];
        int slot = atomicAdd(P.sp.nleave, 1);
        if (slot < P.sp.lcap) P.sp.lrec[slot] = make_int4(i, ch, (item << 8) | dir, r);
        else atomicOr(P.err, NIXB200_ERR_CAPACITY);
      }
      __syncthreads();
      if (tid < 27) {
        int s = 0;
        for (int w = 0; w < THREADS / 32; w++) {
          s += s_warpcnt[w * 27 + tid];
          s_warpcnt[w * 27 + tid] = 0;
        }
        s_dirbase[tid] += s;
      }
      // (next use of s_warpcnt / s_dirbase is behind the barriers of phase 2)
    }
    // the other parity's mover counts are free now (last read before the barrier above)
    for (int t = tid; t < MAXSEG; t += THREADS) s_ccnt[(bpar ^ 1) * MAXSEG + t] = 0;

    // =============================== phase 2: deposit ===============================
    // Per particle and (jy,jz) column, with S0/DS the old weights and the weight differences:
    //   rho[x] += q S1y S1z (S0x+DSx)[x]                                           esirkepov.hpp:155-164
    //   Jx[x]  += -q dx/dt ((S0y+DSy/2) S0z + (S0y/2+DSy/3) DSz) sum_{l<x} DSx[l]            :177-195
    //   Jy[x]  += -q dy/dt sum_{l<jy} DSy[l] ((S0z+DSz/2) S0x[x] + (S0z/2+DSz/3) DSx[x])     :198-216
    //   Jz[x]  += -q dz/dt sum_{l<jz} DSz[l] ((S0x+DSx/2)[x] S0y + (S0x/2+DSx/3)[x] DSy)     :219-237
    const int nmov_w[NWARP] = {s_wcnt[0], s_wcnt[1], s_wcnt[2], s_wcnt[3]};
    int       cbase_e       = 0; // movers of this batch in cells before cc
    const double A = 1.0 / 2, B = 1.0 / 3;
    while (true) {
      const int lo = max(s_pidx[cc], b0) - b0;
      const int hi = min(s_pidx[cc + 1], b1) - b0;

      // ---------- F path: particles that stay in their cell ----------
      if (f_active) {
        for (int p = lo + gf; p < hi; p += GF) {
          if (s_cls[p]) continue;
          const double* sc  = s_scr + (size_t)p * ROW;
          const double2 yv  = *reinterpret_cast<const double2*>(sc + C::PY + 2 * fy);
          const double2 zv  = *reinterpret_cast<const double2*>(sc + C::PZ + 2 * fz);
          const double  cyp = sc[C::CY + fy], czp = sc[C::CZ + fz];
          const double  s0y = yv.x, dsy = yv.y, s0z = zv.x, dsz = zv.y;
          const double  ar = P.q * (s0y + dsy) * (s0z + dsz);
          const double  wx = -((s0y + A * dsy) * s0z + (A * s0y + B * dsy) * dsz) * P.qdxdt[2];
          const double  fyv = -cyp * P.qdxdt[1];
          const double  g0 = fyv * (s0z + A * dsz), g1 = fyv * (A * s0z + B * dsz);
          const double  fzv = -czp * P.qdxdt[0];
          const double  h0 = fzv * s0y, h1 = fzv * dsy;
          const double  k0 = h0 + A * h1, k1 = A * h0 + B * h1;
          double        cpx = 0.0;
#pragma unroll
          for (int s = 0; s < N1; s++) {
            const double2 xv  = *reinterpret_cast<const double2*>(sc + C::PX + 2 * s);
            const double  s0x = xv.x, dsx = xv.y;
            accf[s + 1][0]    = fma(ar, s0x + dsx, accf[s + 1][0]);
            if (s >= 1) accf[s + 1][1] = fma(wx, cpx, accf[s + 1][1]);
            accf[s + 1][2] = fma(g0, s0x, fma(g1, dsx, accf[s + 1][2]));
            accf[s + 1][3] = fma(k0, s0x, fma(k1, dsx, accf[s + 1][3]));
            cpx += dsx;
          }
        }
      }

      // ---------- C path: particles that change cell (block-uniform control flow) ----------
      const int nmc = s_ccnt[bpar * MAXSEG + cc];
      if (nmc > 0) {
        double accc[NS][4];
#pragma unroll
        for (int s = 0; s < NS; s++)
#pragma unroll
          for (int k = 0; k < 4; k++) accc[s][k] = 0.0;
        if (c_active) {
          for (int e = cbase_e + gc; e < cbase_e + nmc; e += GC) {
            // e-th mover of the batch -> (warp list, index)
            int w = 0, r = e;
#pragma unroll
            for (int q = 0; q < NWARP - 1; q++)
              if (w == q && r >= nmov_w[q]) {
                r -= nmov_w[q];
                w = q + 1;
              }
            const int     p   = s_clist[w * 32 + r];
            const double* sc  = s_scr + (size_t)p * ROW;
            double        s0y = 0.0, dsy, s0z = 0.0, dsz;
            if (c_py >= 0) {
              const double2 v = *reinterpret_cast<const double2*>(sc + c_py);
              s0y             = v.x;
              dsy             = v.y;
            } else {
              dsy = sc[c_dy];
            }
            if (c_pz >= 0) {
              const double2 v = *reinterpret_cast<const double2*>(sc + c_pz);
              s0z             = v.x;
              dsz             = v.y;
            } else {
              dsz = sc[c_dz];
            }
            const double cyp = (c_cy >= 0) ? sc[c_cy] : 0.0;
            const double czp = (c_cz >= 0) ? sc[c_cz] : 0.0;
            const double ar = P.q * (s0y + dsy) * (s0z + dsz);
            const double wx = -((s0y + A * dsy) * s0z + (A * s0y + B * dsy) * dsz) * P.qdxdt[2];
            const double fyv = -cyp * P.qdxdt[1];
            const double g0 = fyv * (s0z + A * dsz), g1 = fyv * (A * s0z + B * dsz);
            const double fzv = -czp * P.qdxdt[0];
            const double h0 = fzv * s0y, h1 = fzv * dsy;
            const double k0 = h0 + A * h1, k1 = A * h0 + B * h1;
            double       cpx = 0.0;
#pragma unroll
            for (int s = 0; s < NS; s++) {
              if (s >= 1 && s <= N1) {
                const double2 xv  = *reinterpret_cast<const double2*>(sc + C::PX + 2 * (s - 1));
                const double  s0x = xv.x, dsx = xv.y;
                accc[s][0]        = fma(ar, s0x + dsx, accc[s][0]);
                accc[s][1]        = fma(wx, cpx, accc[s][1]);
                accc[s][2]        = fma(g0, s0x, fma(g1, dsx, accc[s][2]));
                accc[s][3]        = fma(k0, s0x, fma(k1, dsx, accc[s][3]));
                cpx += dsx;
              } else {
                const double dsx = sc[C::OUT + 6 + (s == 0 ? 0 : 1)];
                accc[s][0]       = fma(ar, dsx, accc[s][0]);
                if (s >= 1) accc[s][1] = fma(wx, cpx, accc[s][1]);
                accc[s][2] = fma(g1, dsx, accc[s][2]);
                accc[s][3] = fma(k1, dsx, accc[s][3]);
                cpx += dsx;
              }
            }
          }
        }
        // flush the groups that had work into the window, one group per round (no atomics)
        const int nround = min(nmc, GC);
        for (int r = 0; r < nround; r++) {
          __syncthreads();
          if (c_active && gc == r) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
              double2* w = reinterpret_cast<double2*>(s_win + ((s + rot) % NS) * (NCOL * 4) + col * 4);
              double2  a = w[0], b = w[1];
              a.x += accc[s][0];
              a.y += accc[s][1];
              b.x += accc[s][2];
              b.y += accc[s][3];
              w[0] = a;
              w[1] = b;
            }
          }
        }
        cbase_e += nmc;
      }

      if (cc < ncs && s_pidx[cc + 1] <= b1) {
        retire(); // cell finished: slide the window (block-uniform)
        if (cc >= ncs) break;
      } else {
        break;
      }
    }
    bpar ^= 1;
    __syncthreads(); // scratch may be overwritten by the next batch
  }

  // ---- drain the window --------------------------------------------------------------------------
  for (int s = 0; s < NS - 1; s++) retire();

  // leavers of this work item per direction (scanned over the items of the chunk by k_mig_scan)
  if (tid < 27) P.sp.blockdir[((size_t)ch * g.nitem + item) * 27 + tid] = s_dirbase[tid];
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
  if (a.geo.seg + 1 > MAXSEG) {
    set_error("push segment longer than MAXSEG");
    return 1;
  }
  size_t smem = smem_layout<O>(a.geo.seg).bytes;
  static bool attr_set = false;
  if (!attr_set) {
    NIX_CUDA(cudaFuncSetAttribute(k_push_deposit<O, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  int nblocks = a.geo.nchunk * a.geo.nitem;
  k_push_deposit<O, S><<<nblocks, THREADS, smem, st>>>(*tmap, P);
  NIX_LAUNCHED();
  return 0;
}
} // namespace

size_t push_smem_bytes(const Geo& g)
{
  switch (g.order) {
  case 1: return smem_layout<1>(g.seg).bytes;
  case 2: return smem_layout<2>(g.seg).bytes;
  default: return smem_layout<3>(g.seg).bytes;
  }
}

int launch_push_deposit(const PushArgs& a, const CUtensorMap* tmap, bool strict, cudaStream_t st)
{
  // leaver bookkeeping of this step
  NIX_CUDA(cudaMemsetAsync(a.sp.blockdir, 0, sizeof(int32_t) * (size_t)a.geo.nchunk * a.geo.nitem * 27, st));
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
