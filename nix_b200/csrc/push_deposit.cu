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
//     smem tile, Boris, move, coalesced stores, bin of the new position (-> key + histogram, or a
//     leaver record with its ordered rank inside the item), and the 1-D deposit weights of the
//     particle into a smem scratch.
//   * phase 2, thread per (jy,jz) column of the (O+3)^3 deposit mesh, G groups splitting the
//     particles of the current cell: the Esirkepov current of every particle is accumulated in
//     REGISTERS over all particles of the cell (the GPU analogue of the reference's sorted
//     `reduce_add` path, primitives.hpp:798-809).  The mesh slides along x with the cell index:
//     the column that falls out of the window is reduced over the groups through a small staging
//     buffer and added to the J tile in shared memory (no atomics, fixed order).
//   * the J tile is flushed to global memory once per CTA with red.global.add.f64.
#include "common.cuh"

namespace nixb200
{
namespace
{
constexpr int THREADS = 128;

template <int O>
struct Cfg {
  static constexpr int NW   = O + 2;          // gather stencil width   (interp.hpp)
  static constexpr int NS   = O + 3;          // deposit mesh width     (esirkepov.hpp:326-328)
  static constexpr int NCOL = NS * NS;        // (jy,jz) columns
  static constexpr int G    = THREADS / NCOL; // particle groups: 8 / 5 / 3
  // per-particle scratch, stored field-major ([field][particle], row stride SP) so that the stores of
  // phase 1 (consecutive particles) and the loads of phase 2 (consecutive fields) are conflict free
  static constexpr int X_S0 = 0;              // S0x[1..O+1]        (O+1 values; slots 0 and O+2 are 0)
  static constexpr int X_DS = X_S0 + O + 1;   // DSx[0..NS-1]
  static constexpr int Y_S0 = X_DS + NS, Y_DS = Y_S0 + NS, Y_CP = Y_DS + NS;
  static constexpr int Z_S0 = Y_CP + NS, Z_DS = Z_S0 + NS, Z_CP = Z_DS + NS;
  static constexpr int F    = Z_CP + NS;      // 30 / 38 / 46 doubles per particle
  static constexpr int SP   = THREADS + 1;    // odd row stride
};

struct SmemLayout {
  int    eb_doubles, stage_doubles, scratch_doubles;
  size_t bytes;
};

template <int O>
__host__ __device__ inline SmemLayout smem_layout(int seg)
{
  using C = Cfg<O>;
  SmemLayout L;
  int        ex     = seg + C::NW - 1;
  L.eb_doubles      = ((C::NW * C::NW * ex * 6) + 15) / 16 * 16; // keep 128-byte multiples
  L.stage_doubles   = ((2 * C::G * C::NCOL * 4) + 15) / 16 * 16; // double buffered
  L.scratch_doubles = (C::F * C::SP + 15) / 16 * 16;
  L.bytes = sizeof(double) * ((size_t)L.eb_doubles + L.stage_doubles + L.scratch_doubles) +
            256 * sizeof(int) /* ints */;
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
  constexpr int NS   = C::NS;
  constexpr int NCOL = C::NCOL;
  constexpr int G    = C::G;
  constexpr int SP   = C::SP;
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
  double*   s_stage   = smem_d + L.eb_doubles;
  double*   s_scr     = s_stage + L.stage_doubles;
  int*      s_int     = reinterpret_cast<int*>(s_scr + L.scratch_doubles);
  uint64_t* s_bar     = reinterpret_cast<uint64_t*>(s_int); // 2 ints
  int*      s_pidx    = s_int + 2;                          // [seg+1] <= 66
  int*      s_dirbase = s_int + 72;                         // [27]
  int*      s_warpcnt = s_int + 100;                        // [4][27]

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

  // phase-2 role
  const bool dep_active = tid < G * NCOL;
  const int  grp        = tid / NCOL;
  const int  col        = tid % NCOL;
  const int  jy         = col % NS;
  const int  jz         = col / NS;
  double     acc[NS][4];
#pragma unroll
  for (int s = 0; s < NS; s++)
#pragma unroll
    for (int k = 0; k < 4; k++) acc[s][k] = 0.0;
  int cc  = 0; // current cell (relative to xs) of the sliding deposit window
  int par = 0; // staging buffer parity

  mbar_wait(s_bar, 0);

  // Retire slot 0 of the window: reduce it over the G groups through the staging buffer and add it
  // to J in global memory (every (z,y,x) of the CTA's footprint is retired exactly once, so this
  // is the CTA's single flush of that entry), then slide the window by one cell.
  auto retire = [&]() {
    double* st = s_stage + par * (G * NCOL * 4);
    if (dep_active) {
#pragma unroll
      for (int k = 0; k < 4; k++) st[(grp * NCOL + col) * 4 + k] = acc[0][k];
    }
    __syncthreads();
    for (int t = tid; t < NCOL * 4; t += THREADS) {
      double sum = 0.0;
#pragma unroll
      for (int gg = 0; gg < G; gg++) sum += st[gg * NCOL * 4 + t];
      if (sum != 0.0) {
        int cl = t >> 2, k = t & 3;
        int gy = jy0 + cl % NS, gz = jz0 + cl / NS, gx = jx0 + cc;
        if (gx >= 0 && gx < g.M[2] && gy >= 0 && gy < g.M[1] && gz >= 0 && gz < g.M[0])
          atomicAdd(&ujc[(((size_t)gz * g.M[1] + gy) * g.M[2] + gx) * 4 + k], sum);
      }
    }
    par ^= 1; // the other buffer is free: its readers passed the barrier above
#pragma unroll
    for (int s = 0; s < NS - 1; s++)
#pragma unroll
      for (int k = 0; k < 4; k++) acc[s][k] = acc[s + 1][k];
#pragma unroll
    for (int k = 0; k < 4; k++) acc[NS - 1][k] = 0.0;
    cc++;
  };

  for (int b0 = p_begin; b0 < p_end; b0 += THREADS) {
    const int  b1    = min(b0 + THREADS, p_end);
    const int  i     = b0 + tid;
    const bool valid = i < b1;
    int        dir   = 13;
    double*    scr   = s_scr + tid; // field f of this thread's particle: scr[f * SP]

    // =============================== phase 1: push ===============================
    if (valid) {
      double pos[3], u[3]; // index 0,1,2 = z,y,x
      pos[2] = xu[soa(0, cap, i)];
      pos[1] = xu[soa(1, cap, i)];
      pos[0] = xu[soa(2, cap, i)];
      u[2]   = xu[soa(3, cap, i)];
      u[1]   = xu[soa(4, cap, i)];
      u[0]   = xu[soa(5, cap, i)];

      int    ki[3], sh[3];
      double wi[3][NW], wh[3][NW];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        int ii = digitize(pos[a], c.off[a], g.rdel[a]);
        ki[a]  = ii - g.is_odd;
        int hh = digitize(pos[a], c.hoff[a], g.rdel[a]);
#pragma unroll
        for (int j = 0; j < NW; j++) {
          wi[a][j] = 0.0;
          wh[a][j] = 0.0;
        }
        shape_mc<O, S>(pos[a], add<S>(c.imin[a], mul<S>((double)ki[a], g.del[a])), g.rdel[a], wi[a]);
        shape_mc<O, S>(pos[a], add<S>(c.lo[a], mul<S>((double)hh, g.del[a])), g.rdel[a], wh[a]);
        // interp::shift_weights<O>(hh - ki, wh)  interp.hpp:154-160
        if (hh - ki[a] > 0) {
#pragma unroll
          for (int j = NW - 1; j > 0; j--) wh[a][j] = wh[a][j - 1];
          wh[a][0] = 0.0;
        }
        sh[a] = ii;
      }
      const int  txo       = sh[2] - xs; // cell offset inside the segment
      const bool sorted_ok = (sh[0] == rz) && (sh[1] == ry) && (txo >= 0) && (txo < ncs);
      if (!sorted_ok) atomicOr(P.err, NIXB200_ERR_UNSORTED);
      const int tx = sorted_ok ? txo : 0;

      // ---- gather: 6 components, factorised exactly like interp3d_impl_sorted -----------------
      double rz6[6];
#pragma unroll
      for (int k = 0; k < 6; k++) rz6[k] = 0.0;
#pragma unroll
      for (int jzz = 0; jzz < NW; jzz++) {
        double ry6[6];
#pragma unroll
        for (int k = 0; k < 6; k++) ry6[k] = 0.0;
#pragma unroll
        for (int jyy = 0; jyy < NW; jyy++) {
          double rx6[6];
#pragma unroll
          for (int k = 0; k < 6; k++) rx6[k] = 0.0;
          const double2* pt = reinterpret_cast<const double2*>(s_eb + ((jzz * NW + jyy) * EX + tx) * 6);
#pragma unroll
          for (int jxx = 0; jxx < NW; jxx++) {
            double2 e01 = pt[3 * jxx + 0], e23 = pt[3 * jxx + 1], e45 = pt[3 * jxx + 2];
            rx6[0] = mad<S>(e01.x, wh[2][jxx], rx6[0]); // Ex: half in x
            rx6[1] = mad<S>(e01.y, wi[2][jxx], rx6[1]); // Ey
            rx6[2] = mad<S>(e23.x, wi[2][jxx], rx6[2]); // Ez
            rx6[3] = mad<S>(e23.y, wi[2][jxx], rx6[3]); // Bx
            rx6[4] = mad<S>(e45.x, wh[2][jxx], rx6[4]); // By
            rx6[5] = mad<S>(e45.y, wh[2][jxx], rx6[5]); // Bz
          }
          ry6[0] = mad<S>(rx6[0], wi[1][jyy], ry6[0]);
          ry6[1] = mad<S>(rx6[1], wh[1][jyy], ry6[1]); // Ey: half in y
          ry6[2] = mad<S>(rx6[2], wi[1][jyy], ry6[2]);
          ry6[3] = mad<S>(rx6[3], wh[1][jyy], ry6[3]); // Bx
          ry6[4] = mad<S>(rx6[4], wi[1][jyy], ry6[4]);
          ry6[5] = mad<S>(rx6[5], wh[1][jyy], ry6[5]); // Bz
        }
        rz6[0] = mad<S>(ry6[0], wi[0][jzz], rz6[0]);
        rz6[1] = mad<S>(ry6[1], wi[0][jzz], rz6[1]);
        rz6[2] = mad<S>(ry6[2], wh[0][jzz], rz6[2]); // Ez: half in z
        rz6[3] = mad<S>(ry6[3], wh[0][jzz], rz6[3]); // Bx
        rz6[4] = mad<S>(ry6[4], wh[0][jzz], rz6[4]); // By
        rz6[5] = mad<S>(ry6[5], wi[0][jzz], rz6[5]);
      }
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
#pragma unroll
      for (int a = 0; a < 3; a++) {
        i1[a]   = digitize(pn[a], c.off[a], g.rdel[a]);
        int dd  = (pn[a] >= c.hi[a]) - (pn[a] < c.lo[a]) + 1;
        dcode   = dcode * 3 + dd;
        int sft = (i1[a] - g.is_odd) - ki[a];
        cfl_ok  = cfl_ok && (sft >= -1) && (sft <= 1);
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

      // ---- 1-D deposit weights of this particle -> scratch -----------------------------------
      const bool dep_ok = cfl_ok && sorted_ok;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        double wn[NW];
#pragma unroll
        for (int j = 0; j < NW; j++) wn[j] = 0.0;
        int k1  = i1[a] - g.is_odd;
        int sft = k1 - ki[a];
        shape_mc<O, S>(pn[a], add<S>(c.imin[a], mul<S>((double)k1, g.del[a])), g.rdel[a], wn);
        const int base_s0 = (a == 2) ? C::X_S0 : (a == 1 ? C::Y_S0 : C::Z_S0);
        const int base_ds = (a == 2) ? C::X_DS : (a == 1 ? C::Y_DS : C::Z_DS);
        const int base_cp = (a == 1) ? C::Y_CP : C::Z_CP;
        double    cp      = 0.0;
#pragma unroll
        for (int j = 0; j < NS; j++) {
          // ss[0][.][1..O+1] = old weights; ss[1][.][1+sft..] = new weights (test_esirkepov.cpp:1060-1085)
          double s0 = (j >= 1 && j <= O + 1) ? wi[a][j - 1] : 0.0;
          double vm = (j >= 0 && j <= O) ? wn[j] : 0.0;         // sft = -1 : slot j <- wn[j]
          double v0 = (j >= 1 && j <= O + 1) ? wn[j - 1] : 0.0; // sft =  0
          double vp = (j >= 2 && j <= O + 2) ? wn[j - 2] : 0.0; // sft = +1
          double s1 = (sft == 0) ? v0 : ((sft < 0) ? vm : vp);
          if (!dep_ok) {
            s0 = 0.0;
            s1 = 0.0;
          }
          double ds = s1 - s0; // ds3d, esirkepov.hpp:167-174
          if (a == 2) {
            if (j >= 1 && j <= O + 1) scr[(base_s0 + j - 1) * SP] = s0;
          } else {
            scr[(base_s0 + j) * SP] = s0;
            scr[(base_cp + j) * SP] = cp; // sum of DS over slots < j
          }
          scr[(base_ds + j) * SP] = ds;
          cp += ds;
        }
      }
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

    // =============================== phase 2: deposit ===============================
    // Per particle and (jy,jz) column, with S0/DS the old weights and the weight differences:
    //   rho[x] += q S1y S1z (S0x+DSx)[x]                                           esirkepov.hpp:155-164
    //   Jx[x]  += -q dx/dt ((S0y+DSy/2) S0z + (S0y/2+DSy/3) DSz) sum_{l<x} DSx[l]            :177-195
    //   Jy[x]  += -q dy/dt sum_{l<jy} DSy[l] ((S0z+DSz/2) S0x[x] + (S0z/2+DSz/3) DSx[x])     :198-216
    //   Jz[x]  += -q dz/dt sum_{l<jz} DSz[l] ((S0x+DSx/2)[x] S0y + (S0x/2+DSx/3)[x] DSy)     :219-237
    while (true) {
      const int lo = max(s_pidx[cc], b0);
      const int hi = min(s_pidx[cc + 1], b1);
      if (dep_active) {
        for (int p = lo + grp; p < hi; p += G) {
          const double* sc  = s_scr + (p - b0);
          const double  s0y = sc[(C::Y_S0 + jy) * SP], dsy = sc[(C::Y_DS + jy) * SP], cyp = sc[(C::Y_CP + jy) * SP];
          const double  s0z = sc[(C::Z_S0 + jz) * SP], dsz = sc[(C::Z_DS + jz) * SP], czp = sc[(C::Z_CP + jz) * SP];
          const double  A = 1.0 / 2, B = 1.0 / 3;
          const double  ar = P.q * (s0y + dsy) * (s0z + dsz);
          const double  wx = -((s0y + A * dsy) * s0z + (A * s0y + B * dsy) * dsz) * P.qdxdt[2];
          const double  fy = -cyp * P.qdxdt[1];
          const double  g0 = fy * (s0z + A * dsz), g1 = fy * (A * s0z + B * dsz);
          const double  fz = -czp * P.qdxdt[0];
          const double  h0 = fz * s0y, h1 = fz * dsy;
          const double  k0 = h0 + A * h1, k1 = A * h0 + B * h1;
          double        cpx = 0.0;
#pragma unroll
          for (int s = 0; s < NS; s++) {
            const double dsx = sc[(C::X_DS + s) * SP];
            if (s >= 1 && s <= O + 1) {
              const double s0x = sc[(C::X_S0 + s - 1) * SP];
              acc[s][0] = fma(ar, s0x + dsx, acc[s][0]);
              acc[s][2] = fma(g0, s0x, fma(g1, dsx, acc[s][2]));
              acc[s][3] = fma(k0, s0x, fma(k1, dsx, acc[s][3]));
            } else {
              acc[s][0] = fma(ar, dsx, acc[s][0]);
              acc[s][2] = fma(g1, dsx, acc[s][2]);
              acc[s][3] = fma(k1, dsx, acc[s][3]);
            }
            if (s >= 1) acc[s][1] = fma(wx, cpx, acc[s][1]);
            cpx += dsx;
          }
        }
      }
      if (cc < ncs && s_pidx[cc + 1] <= b1) {
        retire(); // cell finished: slide the window (block-uniform)
        if (cc >= ncs) break;
      } else {
        break;
      }
    }
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
