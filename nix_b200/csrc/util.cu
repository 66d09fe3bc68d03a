// util.cu -- AoS <-> SoA transposes at the C-ABI boundary (sm_100a)
//
// The reference keeps particles as xu[Np][7] (xtensor_particle.hpp:15) and that is the layout of
// Chunk::pack/unpack, of the migration wire format (56 bytes/particle, xtensor_halo3d.hpp:259) and
// of every checkpoint.  On the device particles are SoA; these kernels convert at the boundary.
#include "common.cuh"

namespace nixb200
{
namespace
{
// 32 particles x 7 components staged through shared memory so both sides are coalesced
__global__ void __launch_bounds__(224) k_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soab,
                                                    size_t cap, size_t first, size_t n)
{
  __shared__ double tile[32 * NC];
  for (size_t base = (size_t)blockIdx.x * 32; base < n; base += (size_t)gridDim.x * 32) {
    size_t cnt = (n - base < 32) ? (n - base) : 32;
    int    t   = threadIdx.x;
    if ((size_t)t < cnt * NC) tile[t] = aos[base * NC + t];
    __syncthreads();
    int c = t / 32, p = t % 32;
    if ((size_t)p < cnt) soab[soa(c, cap, first + base + p)] = tile[p * NC + c];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(224) k_soa_to_aos(const double* __restrict__ soab, double* __restrict__ aos,
                                                    size_t cap, size_t first, size_t n)
{
  __shared__ double tile[32 * NC];
  for (size_t base = (size_t)blockIdx.x * 32; base < n; base += (size_t)gridDim.x * 32) {
    size_t cnt = (n - base < 32) ? (n - base) : 32;
    int    t   = threadIdx.x;
    int    c = t / 32, p = t % 32;
    if ((size_t)p < cnt) tile[p * NC + c] = soab[soa(c, cap, first + base + p)];
    __syncthreads();
    if ((size_t)t < cnt * NC) aos[base * NC + t] = tile[t];
    __syncthreads();
  }
}

// interior cells of every chunk <-> a dense [chunk][Nz][Ny][Nx][nc] array (what a host-side field
// solver exchanges with the device: the ghost cells are the halo kernels' business); one thread per
// 16-byte pair of components, rows of the interior are contiguous on both sides
template <bool PACK>
__global__ void __launch_bounds__(256) k_interior(double2* __restrict__ full, double2* __restrict__ dense, int nchunk,
                                                  int Nz, int Ny, int Nx, int nb, int nc2)
{
  const int    My = Ny + 2 * nb, Mx = Nx + 2 * nb, Mz = Nz + 2 * nb;
  const size_t row = (size_t)Nx * nc2; // double2 per interior row
  const size_t n   = (size_t)nchunk * Nz * Ny * row;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t r = t / row, x = t - r * row;
    const int    iy = (int)(r % Ny), iz = (int)((r / Ny) % Nz), k = (int)(r / ((size_t)Ny * Nz));
    const size_t f = ((((size_t)k * Mz + iz + nb) * My + iy + nb) * Mx + nb) * nc2 + x;
    if (PACK) dense[t] = full[f];
    else full[f] = dense[t];
  }
}

// ---- fp32 mode: the C ABI stays the reference's (fp64 AoS particles, fp64 field arrays); the device keeps
//      floats.  Positions are stored relative to the chunk's origin -- a global fp32 coordinate would lose 1e-5
//      of a cell beyond x ~ 100 cells -- and the 64-bit id is kept bit for bit in two words.
__device__ __forceinline__ int chunk_of(const int32_t* __restrict__ cbase, int nchunk, size_t i)
{
  int lo = 0, hi = nchunk;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if ((size_t)cbase[mid] <= i) lo = mid;
    else hi = mid;
  }
  return lo;
}

// origin: [nchunk][3] = (z, y, x) lower corner of every chunk
__global__ void __launch_bounds__(256) k_aos_to_soa_f32(const double* __restrict__ aos, float* __restrict__ soab, size_t cap,
                                                        size_t first, size_t n, const int32_t* __restrict__ cbase,
                                                        int nchunk, const double* __restrict__ origin, double ez, double ey,
                                                        double ex)
{
  // a position strictly inside its chunk must stay inside after rounding to fp32
  auto local = [](double x, double o, double ext) {
    const double r = x - o;
    float        f = (float)r;
    if (r < ext && f >= (float)ext) f = __int_as_float(__float_as_int((float)ext) - 1);
    if (r >= 0.0 && f < 0.f) f = 0.f;
    return f;
  };
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    const double* a  = aos + p * NC;
    const int     ch = chunk_of(cbase, nchunk, first + p);
    const size_t  i  = first + p;
    soab[soa(0, cap, i)] = local(a[0], origin[3 * ch + 2], ex);
    soab[soa(1, cap, i)] = local(a[1], origin[3 * ch + 1], ey);
    soab[soa(2, cap, i)] = local(a[2], origin[3 * ch + 0], ez);
    soab[soa(3, cap, i)] = (float)a[3];
    soab[soa(4, cap, i)] = (float)a[4];
    soab[soa(5, cap, i)] = (float)a[5];
    const long long id   = __double_as_longlong(a[6]);
    soab[soa(6, cap, i)] = __int_as_float((int)(id & 0xffffffffll));
    soab[soa(7, cap, i)] = __int_as_float((int)(id >> 32));
  }
}

__global__ void __launch_bounds__(256) k_soa_to_aos_f32(const float* __restrict__ soab, double* __restrict__ aos, size_t cap,
                                                        size_t first, size_t n, const double* __restrict__ origin3)
{
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    double*      a = aos + p * NC;
    const size_t i = first + p;
    a[0] = origin3[2] + (double)soab[soa(0, cap, i)];
    a[1] = origin3[1] + (double)soab[soa(1, cap, i)];
    a[2] = origin3[0] + (double)soab[soa(2, cap, i)];
    a[3] = (double)soab[soa(3, cap, i)];
    a[4] = (double)soab[soa(4, cap, i)];
    a[5] = (double)soab[soa(5, cap, i)];
    const unsigned lo = (unsigned)__float_as_int(soab[soa(6, cap, i)]);
    const int      hi = __float_as_int(soab[soa(7, cap, i)]);
    a[6] = __longlong_as_double(((long long)hi << 32) | (long long)lo);
  }
}

// fp64 host-layout cells [n][nc] <-> fp32 device cells [n][fc] (fc >= nc: padded E/B cells)
template <bool TO_DEV>
__global__ void __launch_bounds__(256) k_cells_convert(double* __restrict__ h, float* __restrict__ d, size_t ncell, int nc, int fc)
{
  const size_t n = ncell * nc;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t cell = t / nc;
    const int    c    = (int)(t - cell * nc);
    if (TO_DEV) d[cell * fc + c] = (float)h[t];
    else h[t] = (double)d[cell * fc + c];
  }
}

// interior cells of every chunk, fp32 device layout <-> dense fp64 [chunk][Nz][Ny][Nx][nc]
template <bool PACK>
__global__ void __launch_bounds__(256) k_interior_f32(float* __restrict__ full, double* __restrict__ dense, int nchunk, int Nz,
                                                      int Ny, int Nx, int nb, int nc, int fc)
{
  const int    My = Ny + 2 * nb, Mx = Nx + 2 * nb, Mz = Nz + 2 * nb;
  const size_t row = (size_t)Nx * nc;
  const size_t n   = (size_t)nchunk * Nz * Ny * row;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t r = t / row, x = t - r * row;
    const int    ix = (int)(x / nc), c = (int)(x - (size_t)ix * nc);
    const int    iy = (int)(r % Ny), iz = (int)((r / Ny) % Nz), k = (int)(r / ((size_t)Ny * Nz));
    const size_t f = ((((size_t)k * Mz + iz + nb) * My + iy + nb) * Mx + nb + ix) * fc + c;
    if (PACK) dense[t] = (double)full[f];
    else full[f] = (float)dense[t];
  }
}

inline int blocks_for(size_t n)
{
  size_t b = (n + 31) / 32;
  if (b < 1) b = 1;
  if (b > 148 * 32) b = 148 * 32;
  return (int)b;
}
} // namespace

int launch_aos_to_soa(const double* aos, double* soa_base, size_t cap, size_t first, size_t n,
                      cudaStream_t st)
{
  if (n == 0) return 0;
  k_aos_to_soa<<<blocks_for(n), 224, 0, st>>>(aos, soa_base, cap, first, n);
  NIX_LAUNCHED();
  return 0;
}

int launch_soa_to_aos(const double* soa_base, double* aos, size_t cap, size_t first, size_t n,
                      cudaStream_t st)
{
  if (n == 0) return 0;
  k_soa_to_aos<<<blocks_for(n), 224, 0, st>>>(soa_base, aos, cap, first, n);
  NIX_LAUNCHED();
  return 0;
}
static int grid_1d(size_t n)
{
  size_t b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

int launch_aos_to_soa_f32(const double* aos, float* soa_base, size_t cap, size_t first, size_t n, const int32_t* cbase,
                          int nchunk, const double* origin, const double* extent, cudaStream_t st)
{
  if (n == 0) return 0;
  k_aos_to_soa_f32<<<grid_1d(n), 256, 0, st>>>(aos, soa_base, cap, first, n, cbase, nchunk, origin, extent[0], extent[1], extent[2]);
  NIX_LAUNCHED();
  return 0;
}

int launch_soa_to_aos_f32(const float* soa_base, double* aos, size_t cap, size_t first, size_t n, const double* origin3,
                          cudaStream_t st)
{
  if (n == 0) return 0;
  k_soa_to_aos_f32<<<grid_1d(n), 256, 0, st>>>(soa_base, aos, cap, first, n, origin3);
  NIX_LAUNCHED();
  return 0;
}

int launch_cells_convert(bool to_dev, double* host_layout, float* dev_layout, size_t ncell, int nc, int fc, cudaStream_t st)
{
  if (ncell == 0) return 0;
  if (to_dev) k_cells_convert<true><<<grid_1d(ncell * nc), 256, 0, st>>>(host_layout, dev_layout, ncell, nc, fc);
  else k_cells_convert<false><<<grid_1d(ncell * nc), 256, 0, st>>>(host_layout, dev_layout, ncell, nc, fc);
  NIX_LAUNCHED();
  return 0;
}

int launch_interior_f32(bool pack, float* full, double* dense, const Geo& g, int ncomp, int fc, cudaStream_t st)
{
  const size_t n = (size_t)g.nchunk * g.N[0] * g.N[1] * g.N[2] * ncomp;
  if (n == 0) return 0;
  if (pack) k_interior_f32<true><<<grid_1d(n), 256, 0, st>>>(full, dense, g.nchunk, g.N[0], g.N[1], g.N[2], g.nb, ncomp, fc);
  else k_interior_f32<false><<<grid_1d(n), 256, 0, st>>>(full, dense, g.nchunk, g.N[0], g.N[1], g.N[2], g.nb, ncomp, fc);
  NIX_LAUNCHED();
  return 0;
}

namespace
{
__global__ void k_np_from_cbase(const int32_t* __restrict__ cbase, int nch, int64_t* __restrict__ np)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nch) np[k] = (int64_t)cbase[k + 1] - cbase[k];
}
} // namespace

int launch_np_from_cbase(const int32_t* cbase, int nch, int64_t* np, cudaStream_t st)
{
  k_np_from_cbase<<<(nch + 255) / 256, 256, 0, st>>>(cbase, nch, np);
  NIX_LAUNCHED();
  return 0;
}

int launch_interior(bool pack, double* full, double* dense, const Geo& g, int ncomp, cudaStream_t st)
{
  const int    nc2 = ncomp / 2;
  const size_t n   = (size_t)g.nchunk * g.N[0] * g.N[1] * g.N[2] * nc2;
  if (n == 0) return 0;
  size_t b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (pack)
    k_interior<true><<<(int)b, 256, 0, st>>>(reinterpret_cast<double2*>(full), reinterpret_cast<double2*>(dense),
                                             g.nchunk, g.N[0], g.N[1], g.N[2], g.nb, nc2);
  else
    k_interior<false><<<(int)b, 256, 0, st>>>(reinterpret_cast<double2*>(full), reinterpret_cast<double2*>(dense),
                                              g.nchunk, g.N[0], g.N[1], g.N[2], g.nb, nc2);
  NIX_LAUNCHED();
  return 0;
}
} // namespace nixb200
