// diag.cu -- the output side of the path on the device (sm_100a): SURVEY.md section 8f rows N3 and N4
//
//   XtensorPacker3D::pack_field    colocate E/B at the cell centres + decimate   xtensor_packer3d.hpp:62-82,279-302
//   XtensorPacker3D::pack_moment   decimate (block averages) of J / moments      xtensor_packer3d.hpp:84-104,185-230
//   XtensorPacker3D::pack_tracer   the particles with a negative 64-bit id       xtensor_packer3d.hpp:122-140
//   append_moment3d<Order>         (O+1)^3 x 14 moment scatter                   primitives.hpp:896-930
//   XtensorHaloMoment3D            ghost -> neighbour interior, added            xtensor_halo3d.hpp:135-187
//   shape_mc<4>, shape_wt<1..4>    the remaining shape functions                 primitives.hpp:302-495
//
// Without these a diagnostic step downloads every chunk whole (uf + uj + all particles) to run the
// reference's packers on the host; with them only the packed output crosses PCIe.  pack_field /
// pack_moment reproduce the reference's rounding (same products, same summation order, no contraction):
// bit-identical in fp64.  None of this is on the per-step hot path: the kernels are simple, coalesced and
// HBM-bound by construction (one thread per output element or per particle).
#include "common.cuh"

namespace nixb200
{
namespace
{
constexpr int NMOM = 14; // primitives.hpp:900

// ---- shape functions (strict arithmetic, the reference's association order) ---------------------------
__device__ __forceinline__ double M(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double A(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double S(double a, double b) { return __dsub_rn(a, b); }

__device__ void shape_mc_any(int order, double x, double X, double rdx, double* s)
{
  const double d = M(S(x, X), rdx);
  if (order == 1) {
    s[0] = S(1.0, d);
    s[1] = d;
  } else if (order == 2) {
    const double w1 = S(0.5, d), w2 = A(0.5, d);
    s[0] = M(M(0.50, w1), w1);
    s[1] = S(0.75, M(d, d));
    s[2] = M(M(0.50, w2), w2);
  } else if (order == 3) {
    const double a = 1 / 6.0, w1 = d, w2 = S(1.0, d);
    const double w12 = M(w1, w1), w22 = M(w2, w2), w13 = M(w12, w1), w23 = M(w22, w2);
    s[0] = M(a, w23);
    s[1] = M(a, A(S(4.0, M(6.0, w12)), M(3.0, w13)));
    s[2] = M(a, A(S(4.0, M(6.0, w22)), M(3.0, w23)));
    s[3] = M(a, w13);
  } else { // order 4, primitives.hpp:302-331
    const double a = 1 / 384.0, b = 1 / 96.0, c = 115 / 192.0, e = 1 / 8.0;
    const double p1 = A(1.0, d), m1 = S(1.0, d), p2 = A(1.0, M(d, 2.0)), m2 = S(1.0, M(d, 2.0));
    const double d2 = M(d, d);
    const double p12 = M(p1, p1), m12 = M(m1, m1), p13 = M(p12, p1), m13 = M(m12, m1), p14 = M(p13, p1), m14 = M(m13, m1);
    s[0] = M(a, M(M(M(m2, m2), m2), m2));
    s[1] = M(b, S(A(S(A(55.0, M(20.0, p1)), M(120.0, p12)), M(80.0, p13)), M(16.0, p14)));
    s[2] = A(c, M(M(e, d2), S(M(2.0, d2), 5.0)));
    s[3] = M(b, S(A(S(A(55.0, M(20.0, m1)), M(120.0, m12)), M(80.0, m13)), M(16.0, m14)));
    s[4] = M(a, M(M(M(p2, p2), p2), p2));
  }
}

// shape functions of the WT scheme: three branches selected by where delta sits, blended by 0/1 masks
__device__ void shape_wt_any(int order, double x, double X, double rdx, double dt, double rdt, double* s)
{
  const double d = M(S(x, X), rdx);
  if (order == 1) {
    const double v = M(M(0.25, rdt), S(A(1.0, M(2.0, dt)), M(2.0, d)));
    const double ss = fmin(1.0, fmax(0.0, v));
    s[0] = ss;
    s[1] = S(1.0, ss);
    return;
  }
  const bool   odd = order & 1;
  const double lo = odd ? S(0.5, dt) : -dt, hi = odd ? A(0.5, dt) : dt;
  const double t1 = (d < lo) ? 1.0 : 0.0, t2 = S(1.0, t1), t3 = (d < hi) ? 1.0 : 0.0, t4 = S(1.0, t3);
  double       P[5] = {0, 0, 0, 0, 0}, Q[5] = {0, 0, 0, 0, 0}, R[5] = {0, 0, 0, 0, 0};
  if (order == 2) {
    const double w0 = fabs(d), w1 = S(dt, d), w2 = A(dt, d);
    P[0] = w0, P[1] = S(1.0, w0), P[2] = 0.0;
    Q[0] = M(M(M(0.25, rdt), w1), w1);
    Q[1] = M(M(0.50, rdt), S(M(dt, S(2.0, dt)), M(w0, w0)));
    Q[2] = M(M(M(0.25, rdt), w2), w2);
    R[0] = P[2], R[1] = P[1], R[2] = P[0];
  } else if (order == 3) {
    const double a = 1 / 96.0, b = 1 / 24.0, c = 1 / 12.0, adt = M(a, rdt);
    const double w0 = d, w1 = S(1.0, d), w3 = S(1.0, M(2.0, d)), w4 = A(1.0, M(2.0, d));
    const double w5 = A(M(2.0, dt), w3), w6 = S(M(2.0, dt), w3), w7 = S(3.0, M(2.0, d));
    const double w02 = M(w0, w0), w12 = M(w1, w1), w32 = M(w3, w3), w33 = M(w32, w3), w42 = M(w4, w4);
    const double w53 = M(M(w5, w5), w5), w63 = M(M(w6, w6), w6), w72 = M(w7, w7);
    const double dt2 = M(dt, dt), dt3 = M(dt2, dt), dt24 = M(4.0, dt2);
    const double so = M(adt, S(M(-8.0, dt3), M(M(6.0, dt), w32)));
    const double se = M(adt, S(M(M(-36.0, dt2), w3), M(3.0, w33)));
    P[0] = M(b, A(dt24, M(3.0, w32)));
    P[1] = M(c, S(S(9.0, dt24), M(12.0, w02)));
    P[2] = M(b, A(dt24, M(3.0, w42)));
    Q[0] = M(adt, w53);
    Q[1] = A(A(so, se), w1);
    Q[2] = A(S(so, se), w0);
    Q[3] = M(adt, w63);
    R[1] = M(b, A(dt24, M(3.0, w72)));
    R[2] = M(c, S(S(9.0, dt24), M(12.0, w12)));
    R[3] = M(b, A(dt24, M(3.0, w32)));
  } else {
    const double a = 1 / 48.0, b = 1 / 24.0, c = 1 / 12.0, e = 1 / 6.0, adt = M(a, rdt), bdt = M(b, rdt), cdt = M(c, rdt);
    const double w0 = fabs(d), w1 = S(1.0, w0), w2 = S(1.0, d), w3 = A(1.0, d), w4 = S(dt, d), w5 = A(dt, d);
    const double w02 = M(w0, w0), w03 = M(w02, w0), w04 = M(w03, w0), w12 = M(w1, w1), w13 = M(w12, w1);
    const double w23 = M(M(w2, w2), w2), w33 = M(M(w3, w3), w3);
    const double w44 = M(M(M(w4, w4), w4), w4), w54 = M(M(M(w5, w5), w5), w5);
    const double dt2 = M(dt, dt), dt3 = M(dt2, dt), dt4 = M(dt3, dt);
    const double ss1 = S(S(-dt4, M(M(6.0, w02), dt2)), w04);
    const double ss2 = A(A(A(S(M(3.0, dt4), M(8.0, dt3)), M(M(18.0, w02), dt2)), M(S(16.0, M(24.0, w02)), dt)), M(3.0, w04));
    P[0] = M(M(e, w0), A(w02, dt2));
    P[1] = M(e, A(A(S(4.0, M(6.0, w12)), M(3.0, w13)), M(S(1.0, M(3.0, w0)), dt2)));
    P[2] = M(e, S(A(S(4.0, M(6.0, w02)), M(3.0, w03)), M(S(2.0, M(3.0, w0)), dt2)));
    P[3] = M(M(e, w1), A(w12, dt2));
    Q[0] = M(adt, w44);
    Q[1] = M(cdt, A(A(ss1, M(M(2.0, dt3), w3)), M(M(2.0, dt), A(M(-6.0, d), w33))));
    Q[2] = M(bdt, ss2);
    Q[3] = M(cdt, A(A(ss1, M(M(2.0, dt3), w2)), M(M(2.0, dt), A(M(6.0, d), w23))));
    Q[4] = M(adt, w54);
    for (int j = 0; j < 5; j++) R[j] = P[4 - j];
  }
  for (int j = 0; j <= order; j++) s[j] = A(A(M(P[j], t1), M(M(Q[j], t2), t3)), M(R[j], t4));
}

__global__ void k_shape_eval(int kind, int order, int n, const double* __restrict__ x, const double* __restrict__ X,
                             double rdx, double dt, double rdt, double* __restrict__ out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s[5];
  if (kind == 0) shape_mc_any(order, x[i], X[i], rdx, s);
  else shape_wt_any(order, x[i], X[i], rdx, dt, rdt, s);
  for (int j = 0; j <= order; j++) out[(size_t)i * (order + 1) + j] = s[j];
}

// ---- packers ---------------------------------------------------------------------------------------------
struct PackGeo {
  int M[3], N[3], nb;
  int sz, sy, sx;   // decimated size
  int bz, by, bx;   // block size
  int nc, fc;       // components packed, words per cell on the device
  double factor;
};

// colocated value of component ic at array cell (iz,iy,ix) (colocate_field_3d, xtensor_packer3d.hpp:279-302)
template <typename T>
__device__ __forceinline__ double colocated(const T* __restrict__ u, const PackGeo& g, int iz, int iy, int ix, int ic)
{
  const size_t sx = g.fc, sy = (size_t)g.M[2] * g.fc, sz = (size_t)g.M[1] * g.M[2] * g.fc;
  const T*     p  = u + iz * sz + iy * sy + ix * sx + ic;
  switch (ic) {
  case 0: return M(0.50, A((double)p[0], (double)p[sx]));
  case 1: return M(0.50, A((double)p[0], (double)p[sy]));
  case 2: return M(0.50, A((double)p[0], (double)p[sz]));
  case 3: return M(0.25, A(A(A((double)p[0], (double)p[sz + sy]), (double)p[sy]), (double)p[sz]));
  case 4: return M(0.25, A(A(A((double)p[0], (double)p[sz + sx]), (double)p[sz]), (double)p[sx]));
  default: return M(0.25, A(A(A((double)p[0], (double)p[sy + sx]), (double)p[sx]), (double)p[sy]));
  }
}

// one thread per output element; the block offsets are walked in the reference's order (decimate_field
// accumulates z += factor * y block offset by block offset, xtensor_packer3d.hpp:205-224)
template <typename T, bool COLOCATE>
__global__ void __launch_bounds__(256) k_pack_grid(PackGeo g, const T* __restrict__ u, double* __restrict__ out)
{
  const int n = g.sz * g.sy * g.sx * g.nc;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int ic = t % g.nc, r = t / g.nc;
    const int jx = r % g.sx, jy = (r / g.sx) % g.sy, jz = r / (g.sx * g.sy);
    double    z  = 0.0;
    for (int kz = 0; kz < g.bz; kz++)
      for (int ky = 0; ky < g.by; ky++)
        for (int kx = 0; kx < g.bx; kx++) {
          const int iz = g.nb + jz * g.bz + kz, iy = g.nb + jy * g.by + ky, ix = g.nb + jx * g.bx + kx;
          double    y;
          if (COLOCATE) y = colocated<T>(u, g, iz, iy, ix, ic);
          else y = (double)u[(((size_t)iz * g.M[1] + iy) * g.M[2] + ix) * g.fc + ic];
          z = A(z, M(g.factor, y));
        }
    out[t] = z;
  }
}

// tracers of one chunk: ordered compaction by ONE block (block-wide scan of the flags, tile by tile)
template <typename T>
__global__ void __launch_bounds__(1024) k_pack_tracer(const T* __restrict__ xu, size_t cap, int first, int np,
                                                      const double* __restrict__ origin3, double* __restrict__ out,
                                                      int max_out, int* __restrict__ count)
{
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < np; t0 += blockDim.x) {
    const int i    = t0 + threadIdx.x;
    bool      flag = false;
    long long id   = 0;
    if (i < np) {
      if constexpr (sizeof(T) == 8) id = __double_as_longlong(xu[soa(6, cap, first + i)]);
      else
        id = ((long long)__float_as_int(xu[soa(7, cap, first + i)]) << 32) |
             (long long)(unsigned)__float_as_int(xu[soa(6, cap, first + i)]);
      flag = id < 0;
    }
    const unsigned m   = __ballot_sync(0xffffffffu, flag);
    const int      rk  = __popc(m & ((1u << lane) - 1));
    if (lane == 0) s_warp[w] = __popc(m);
    __syncthreads();
    int before = s_base;
    for (int q = 0; q < w; q++) before += s_warp[q];
    if (flag) {
      const int o = before + rk;
      if (o < max_out) {
        double* d = out + (size_t)o * NC;
        if constexpr (sizeof(T) == 8) {
          for (int c = 0; c < 6; c++) d[c] = xu[soa(c, cap, first + i)];
        } else {
          for (int c = 0; c < 3; c++) d[c] = origin3[2 - c] + (double)xu[soa(c, cap, first + i)];
          for (int c = 3; c < 6; c++) d[c] = (double)xu[soa(c, cap, first + i)];
        }
        d[6] = __longlong_as_double(id);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int q = 0; q < (int)(blockDim.x >> 5); q++) tot += s_warp[q];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = s_base;
}

// ---- moments ---------------------------------------------------------------------------------------------
// One warp per bin of one chunk (the container is cell-sorted: a bin is one contiguous particle range);
// lane = particle.  Node by node the lanes evaluate w * mom[k], the warp sums the 32 values with shuffles
// and lane 0 adds the sum to the chunk's moment array: one atomic per (node, moment) and 32 particles.
// Composition (ours, as the push: the reference ships the scatter, append_moment3d, not the loop):
//   mom = { m;  m u_i/gamma;  m gamma c^2;  m u_i c;  m u_i u_j / gamma (xx, yy, zz, xy, yz, zx) }
template <int O, typename T>
__global__ void __launch_bounds__(256) k_moment(Geo g, const ChunkGeo* __restrict__ cg, SpeciesDev sp, int is, int ns,
                                                double mass, T* __restrict__ um)
{
  constexpr int N1 = O + 1;
  const int     lane = threadIdx.x & 31;
  const int     warps_per_block = blockDim.x >> 5;
  const long long nbin = (long long)g.nchunk * g.ncell;
  const T* __restrict__ xu = reinterpret_cast<const T*>(sp.xu);
  const T del[3] = {(T)g.del[0], (T)g.del[1], (T)g.del[2]}, rdel[3] = {(T)g.rdel[0], (T)g.rdel[1], (T)g.rdel[2]};
  const T cc = (T)g.cc, rc = (T)g.rc, m = (T)mass;
  for (long long b = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); b < nbin;
       b += (long long)gridDim.x * warps_per_block) {
    const int ch = (int)(b / g.ncell), cell = (int)(b - (long long)ch * g.ncell);
    const int i0 = sp.start[(size_t)b * LANES], i1 = sp.start[(size_t)(b + 1) * LANES];
    if (i1 == i0) continue;
    const ChunkGeo& c = cg[ch];
    const int bz = cell / (g.R[1] * g.R[2]), by = (cell / g.R[2]) % g.R[1], bx = cell % g.R[2];
    const int base[3] = {bz - g.is_odd - g.half + g.nb, by - g.is_odd - g.half + g.nb, bx - g.is_odd - g.half + g.nb};
    T* umc = um + (size_t)ch * g.M[0] * g.M[1] * g.M[2] * ns * NMOM;
    for (int p0 = i0; p0 < i1; p0 += 32) {
      const int  i     = p0 + lane;
      const bool valid = i < i1;
      T          w[3][N1], mom[NMOM];
#pragma unroll
      for (int k = 0; k < NMOM; k++) mom[k] = T(0);
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int j = 0; j < N1; j++) w[a][j] = T(0);
      if (valid) {
        const T pos[3] = {xu[soa(2, sp.cap, i)], xu[soa(1, sp.cap, i)], xu[soa(0, sp.cap, i)]};
        const int bin[3] = {bz, by, bx};
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const int ki = bin[a] - g.is_odd;
          const T   Xn = add<true>((T)c.imin[a], mul<true>((T)ki, del[a]));
          const T   d  = mul<true>(sub<true>(pos[a], Xn), rdel[a]);
          if constexpr (O == 1) {
            w[a][0] = sub<true>(T(1), d), w[a][1] = d;
          } else if constexpr (O == 2) {
            const T w1 = sub<true>(T(0.5), d), w2 = add<true>(T(0.5), d);
            w[a][0] = mul<true>(mul<true>(T(0.5), w1), w1);
            w[a][1] = sub<true>(T(0.75), mul<true>(d, d));
            w[a][2] = mul<true>(mul<true>(T(0.5), w2), w2);
          } else {
            const T a6 = T(1 / 6.0), w1 = d, w2 = sub<true>(T(1), d);
            const T w12 = mul<true>(w1, w1), w22 = mul<true>(w2, w2), w13 = mul<true>(w12, w1), w23 = mul<true>(w22, w2);
            w[a][0] = mul<true>(a6, w23);
            w[a][1] = mul<true>(a6, add<true>(sub<true>(T(4), mul<true>(T(6), w12)), mul<true>(T(3), w13)));
            w[a][2] = mul<true>(a6, add<true>(sub<true>(T(4), mul<true>(T(6), w22)), mul<true>(T(3), w23)));
            w[a][3] = mul<true>(a6, w13);
          }
        }
        const T ux = xu[soa(3, sp.cap, i)], uy = xu[soa(4, sp.cap, i)], uz = xu[soa(5, sp.cap, i)];
        const T uu = ux * ux + uy * uy + uz * uz;
        const T gam = sqrt_<true>(T(1) + uu * rc * rc); // lorentz_factor, primitives.hpp:158-161
        mom[0] = m, mom[1] = m * ux / gam, mom[2] = m * uy / gam, mom[3] = m * uz / gam, mom[4] = m * gam * cc * cc;
        mom[5] = m * ux * cc, mom[6] = m * uy * cc, mom[7] = m * uz * cc;
        mom[8] = m * ux * ux / gam, mom[9] = m * uy * uy / gam, mom[10] = m * uz * uz / gam;
        mom[11] = m * ux * uy / gam, mom[12] = m * uy * uz / gam, mom[13] = m * uz * ux / gam;
      }
#pragma unroll
      for (int jz = 0; jz < N1; jz++)
#pragma unroll
        for (int jy = 0; jy < N1; jy++)
#pragma unroll
          for (int jx = 0; jx < N1; jx++) {
            const T ww = (w[0][jz] * w[1][jy]) * w[2][jx];
            T*      dst = umc + ((((size_t)(base[0] + jz) * g.M[1] + (base[1] + jy)) * g.M[2] + (base[2] + jx)) * ns + is) * NMOM;
#pragma unroll
            for (int k = 0; k < NMOM; k++) {
              T v = ww * mom[k];
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
              if (lane == 0 && v != T(0)) atomicAdd(dst + k, v);
            }
          }
    }
  }
}

PackGeo make_pack_geo(const Geo& g, int decimate, int nc, int fc)
{
  PackGeo p;
  for (int a = 0; a < 3; a++) {
    p.M[a] = g.M[a];
    p.N[a] = g.N[a];
  }
  p.nb = g.nb;
  auto dsize = [&](int n) { return (n <= decimate) ? 1 : n / decimate; }; // decimate_size, xtensor_packer3d.hpp:232-241
  p.sz = dsize(g.N[0]), p.sy = dsize(g.N[1]), p.sx = dsize(g.N[2]);
  p.bz = g.N[0] / p.sz, p.by = g.N[1] / p.sy, p.bx = g.N[2] / p.sx;
  p.nc = nc, p.fc = fc;
  p.factor = 1.0 / (p.bz * p.by * p.bx);
  return p;
}
} // namespace

int pack_count(const Geo& g, int decimate, int nc)
{
  PackGeo p = make_pack_geo(g, decimate, nc, nc);
  return p.sz * p.sy * p.sx * nc;
}

// chunk_ptr: first cell of the chunk's array on the device; out: device fp64 [sz][sy][sx][nc]
int launch_pack_grid(const Geo& g, const void* chunk_ptr, bool colocate, int decimate, int nc, int fc, double* out,
                     cudaStream_t st, bool fp32)
{
  if (decimate < 1) {
    set_error("decimate must be >= 1");
    return 1;
  }
  PackGeo   p = make_pack_geo(g, decimate, nc, fc);
  const int n = p.sz * p.sy * p.sx * nc;
  const int blocks = std::max(1, std::min((n + 255) / 256, 148 * 8));
  if (fp32) {
    if (colocate) k_pack_grid<float, true><<<blocks, 256, 0, st>>>(p, (const float*)chunk_ptr, out);
    else k_pack_grid<float, false><<<blocks, 256, 0, st>>>(p, (const float*)chunk_ptr, out);
  } else {
    if (colocate) k_pack_grid<double, true><<<blocks, 256, 0, st>>>(p, (const double*)chunk_ptr, out);
    else k_pack_grid<double, false><<<blocks, 256, 0, st>>>(p, (const double*)chunk_ptr, out);
  }
  NIX_LAUNCHED();
  return 0;
}

int launch_pack_tracer(const SpeciesDev& sp, int first, int np, const double* origin3, double* out, int max_out,
                       int* count_dev, cudaStream_t st, bool fp32)
{
  if (fp32) k_pack_tracer<float><<<1, 1024, 0, st>>>((const float*)sp.xu, (size_t)sp.cap, first, np, origin3, out, max_out, count_dev);
  else k_pack_tracer<double><<<1, 1024, 0, st>>>((const double*)sp.xu, (size_t)sp.cap, first, np, origin3, out, max_out, count_dev);
  NIX_LAUNCHED();
  return 0;
}

int launch_moment(const Geo& g, const ChunkGeo* cg, const SpeciesDev& sp, int is, int ns, void* um, cudaStream_t st,
                  bool fp32)
{
  const int blocks = 148 * 8;
#define NIX_MOM(O)                                                                                                       \
  if (fp32) k_moment<O, float><<<blocks, 256, 0, st>>>(g, cg, sp, is, ns, sp.m, (float*)um);                            \
  else k_moment<O, double><<<blocks, 256, 0, st>>>(g, cg, sp, is, ns, sp.m, (double*)um)
  switch (g.order) {
  case 1: NIX_MOM(1); break;
  case 2: NIX_MOM(2); break;
  default: NIX_MOM(3); break;
  }
#undef NIX_MOM
  NIX_LAUNCHED();
  return 0;
}

int launch_shape_eval(int kind, int order, int n, const double* x, const double* X, double rdx, double dt, double rdt,
                      double* out, cudaStream_t st)
{
  if (n <= 0) return 0;
  k_shape_eval<<<(n + 255) / 256, 256, 0, st>>>(kind, order, n, x, X, rdx, dt, rdt, out);
  NIX_LAUNCHED();
  return 0;
}
} // namespace nixb200
