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
