// fdtd.cu -- Yee FDTD field update on the device-resident grid (sm_100a): SURVEY.md section 8f, row N1
//
// The reference ships no field solver: Application::push() is an empty virtual
// (application.hpp:343-346) and the Maxwell update belongs to the downstream application.  What the
// reference does fix is the staggering of the six components of uf -- through the way it colocates them
// (xtensor_packer3d.hpp:279-302) and through the half-grid gather of the push -- and that J is staggered
// like E (esirkepov.hpp:177-237).  With cell centres at (i + 1/2) dx and edges at i dx:
//     Ex (c,c,e)  Ey (c,e,c)  Ez (e,c,c)      Bx (e,e,c)  By (e,c,e)  Bz (c,e,e)       (z,y,x)
// Keeping this update on the device removes the only per-step host round trip of a nix application
// (interior J down, interior E/B up; 168 MB per step and GPU at the benchmark size).
//
//     k_push_bfd   B -= c dt curl E    on the interior plus `ext` ghost layers (reads E one cell below)
//     k_push_efd   E += c dt curl B - cfj dt J   on the interior (reads B one cell above)
//     k_field_energy   sum E^2, sum B^2 per chunk (history diagnostic; fixed summation order)
//
// HBM-bound and tiny next to the particle kernels (48 B read + 24 B written per cell, ~1/128 of the
// particle bytes at the benchmark size): one thread per cell, rows along x coalesced, neighbours served
// by L1/L2.  Every product and sum is an explicit round-to-nearest operation in the order of
// oracle/field_solver.c, so the result is bit-identical to the CPU restatement.
#include "common.cuh"

namespace nixb200
{
namespace
{
template <typename T>
struct FdtdGeo {
  int M[3], N[3], nb;
  T   cz, cy, cx, cj;
};

// explicit round-to-nearest operations in the oracle's order (no contraction) in both real types
template <typename T> __device__ __forceinline__ T m_(T a, T b) { return mul<true>(a, b); }
template <typename T> __device__ __forceinline__ T s_(T a, T b) { return sub<true>(a, b); }
template <typename T> __device__ __forceinline__ T a_(T a, T b) { return add<true>(a, b); }

// region = [nb-ext, nb+N-1+ext] on every axis; grid.y = chunk
template <typename T>
__global__ void __launch_bounds__(256) k_push_bfd(FdtdGeo<T> g, T* __restrict__ uf, int ext)
{
  constexpr int FC = field_stride<T>();
  const int    ez = g.N[0] + 2 * ext, ey = g.N[1] + 2 * ext, ex = g.N[2] + 2 * ext;
  const int    n  = ez * ey * ex;
  const size_t sy = (size_t)g.M[2] * FC, sz = (size_t)g.M[1] * g.M[2] * FC;
  T*           u  = uf + (size_t)blockIdx.y * g.M[0] * sz;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int ix = t % ex + g.nb - ext, iy = (t / ex) % ey + g.nb - ext, iz = t / (ex * ey) + g.nb - ext;
    T*        p  = u + iz * sz + iy * sy + (size_t)ix * FC;
    const T exx = p[0], eyy = p[1], ezz = p[2];
    const T ez_ym = p[2 - (ptrdiff_t)sy], ex_ym = p[0 - (ptrdiff_t)sy];
    const T ey_zm = p[1 - (ptrdiff_t)sz], ex_zm = p[0 - (ptrdiff_t)sz];
    const T ez_xm = p[2 - FC], ey_xm = p[1 - FC];
    p[3] = s_(p[3], s_(m_(g.cy, s_(ezz, ez_ym)), m_(g.cz, s_(eyy, ey_zm))));
    p[4] = s_(p[4], s_(m_(g.cz, s_(exx, ex_zm)), m_(g.cx, s_(ezz, ez_xm))));
    p[5] = s_(p[5], s_(m_(g.cx, s_(eyy, ey_xm)), m_(g.cy, s_(exx, ex_ym))));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_push_efd(FdtdGeo<T> g, T* __restrict__ uf, const T* __restrict__ uj)
{
  constexpr int FC = field_stride<T>();
  const int    n  = g.N[0] * g.N[1] * g.N[2];
  const size_t sy = (size_t)g.M[2] * FC, sz = (size_t)g.M[1] * g.M[2] * FC;
  T*           u  = uf + (size_t)blockIdx.y * g.M[0] * sz;
  const T*     j  = uj + (size_t)blockIdx.y * g.M[0] * g.M[1] * g.M[2] * 4;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int ix = t % g.N[2] + g.nb, iy = (t / g.N[2]) % g.N[1] + g.nb, iz = t / (g.N[2] * g.N[1]) + g.nb;
    T*       p  = u + iz * sz + iy * sy + (size_t)ix * FC;
    const T* pj = j + (((size_t)iz * g.M[1] + iy) * g.M[2] + ix) * 4;
    const T bx = p[3], by = p[4], bz = p[5];
    const T bz_yp = p[5 + sy], bx_yp = p[3 + sy];
    const T by_zp = p[4 + sz], bx_zp = p[3 + sz];
    const T bz_xp = p[5 + FC], by_xp = p[4 + FC];
    p[0] = s_(a_(p[0], s_(m_(g.cy, s_(bz_yp, bz)), m_(g.cz, s_(by_zp, by)))), m_(g.cj, pj[1]));
    p[1] = s_(a_(p[1], s_(m_(g.cz, s_(bx_zp, bx)), m_(g.cx, s_(bz_xp, bz)))), m_(g.cj, pj[2]));
    p[2] = s_(a_(p[2], s_(m_(g.cx, s_(by_xp, by)), m_(g.cy, s_(bx_yp, bx)))), m_(g.cj, pj[3]));
  }
}

// one block per chunk; per-thread partial sums over a fixed stride, then a fixed shared-memory tree:
// the result does not depend on scheduling
template <typename T>
__global__ void __launch_bounds__(256) k_field_energy(FdtdGeo<T> g, const T* __restrict__ uf, double* __restrict__ out)
{
  constexpr int FC = field_stride<T>();
  __shared__ double s_e[256], s_b[256];
  const int    n  = g.N[0] * g.N[1] * g.N[2];
  const size_t sy = (size_t)g.M[2] * FC, sz = (size_t)g.M[1] * g.M[2] * FC;
  const T*     u  = uf + (size_t)blockIdx.x * g.M[0] * sz;
  double       e2 = 0.0, b2 = 0.0; // (the diagnostic sums in fp64 in both modes)
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int ix = t % g.N[2] + g.nb, iy = (t / g.N[2]) % g.N[1] + g.nb, iz = t / (g.N[2] * g.N[1]) + g.nb;
    const T*  p  = u + iz * sz + iy * sy + (size_t)ix * FC;
    const double e0 = p[0], e1 = p[1], e2_ = p[2], b0 = p[3], b1 = p[4], b2_ = p[5];
    e2 += e0 * e0 + e1 * e1 + e2_ * e2_;
    b2 += b0 * b0 + b1 * b1 + b2_ * b2_;
  }
  s_e[threadIdx.x] = e2;
  s_b[threadIdx.x] = b2;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
      s_e[threadIdx.x] += s_e[threadIdx.x + w];
      s_b[threadIdx.x] += s_b[threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x + 0] = s_e[0];
    out[2 * blockIdx.x + 1] = s_b[0];
  }
}

template <typename T>
FdtdGeo<T> make_geo(const Geo& g, double delt, double cfj)
{
  FdtdGeo<T> f;
  for (int a = 0; a < 3; a++) {
    f.M[a] = g.M[a];
    f.N[a] = g.N[a];
  }
  f.nb = g.nb;
  f.cz = (T)(g.cc * delt / g.del[0]); // the oracle's expressions (oracle/field_solver.c)
  f.cy = (T)(g.cc * delt / g.del[1]);
  f.cx = (T)(g.cc * delt / g.del[2]);
  f.cj = (T)(cfj * delt);
  return f;
}
} // namespace

int launch_push_bfd(const Geo& g, void* uf, double delt, int ext, cudaStream_t st, bool fp32)
{
  if (ext < 0 || ext >= g.nb) {
    set_error("push_bfd: ext must be in [0, nb)");
    return 1;
  }
  const int n = (g.N[0] + 2 * ext) * (g.N[1] + 2 * ext) * (g.N[2] + 2 * ext);
  dim3      grid((n + 255) / 256, g.nchunk);
  if (fp32) k_push_bfd<float><<<grid, 256, 0, st>>>(make_geo<float>(g, delt, 0.0), (float*)uf, ext);
  else k_push_bfd<double><<<grid, 256, 0, st>>>(make_geo<double>(g, delt, 0.0), (double*)uf, ext);
  NIX_LAUNCHED();
  return 0;
}

int launch_push_efd(const Geo& g, void* uf, const void* uj, double delt, double cfj, cudaStream_t st, bool fp32)
{
  const int n = g.N[0] * g.N[1] * g.N[2];
  dim3      grid((n + 255) / 256, g.nchunk);
  if (fp32) k_push_efd<float><<<grid, 256, 0, st>>>(make_geo<float>(g, delt, cfj), (float*)uf, (const float*)uj);
  else k_push_efd<double><<<grid, 256, 0, st>>>(make_geo<double>(g, delt, cfj), (double*)uf, (const double*)uj);
  NIX_LAUNCHED();
  return 0;
}

int launch_field_energy(const Geo& g, const void* uf, double* out, cudaStream_t st, bool fp32)
{
  if (fp32) k_field_energy<float><<<g.nchunk, 256, 0, st>>>(make_geo<float>(g, 0.0, 0.0), (const float*)uf, out);
  else k_field_energy<double><<<g.nchunk, 256, 0, st>>>(make_geo<double>(g, 0.0, 0.0), (const double*)uf, out);
  NIX_LAUNCHED();
  return 0;
}
} // namespace nixb200
