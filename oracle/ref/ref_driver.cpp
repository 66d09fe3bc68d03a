// ref_driver.cpp -- implements oracle/nix_oracle.h by calling the REFERENCE's own templates and
// classes, compiled from where they lie under /root/reference (never copied into this repo).
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/ (git-ignored).
//
// What is the reference's and what is ours:
//   * every numerical primitive, the particle container, the halo classes and the pack/unpack
//     orchestration are the reference's (primitives.hpp, interp.hpp, esirkepov.hpp,
//     xtensor_particle.hpp, xtensor_halo3d.hpp, chunk.hpp/.cpp);
//   * the composed per-particle step (push_deposit below) is OURS, because the reference tree does
//     not contain it (Application::push() is an empty virtual, application.hpp:343-346).  It
//     follows the per-particle call order of unittest/test_esirkepov.cpp:1030-1106 and the
//     staggering documented by xtensor_packer3d.hpp:279-302.  oracle/nix_oracle.c restates the very
//     same composition in plain C; DESIGN.md section 2 documents every choice.
#include "chunk.hpp"
#include "esirkepov.hpp"
#include "interp.hpp"
#include "primitives.hpp"
#include "xtensor_halo3d.hpp"
#include "xtensor_packer3d.hpp"

#include "../nix_oracle.h"

using namespace nix;

namespace
{
class RefChunk : public Chunk
{
public:
  int                     order;
  int                     Ns;
  xt::xtensor<float64, 4> uf;
  xt::xtensor<float64, 4> uj;
  xt::xtensor<float64, 5> um; // [Mz][My][Mx][Ns][14]
  ParticleVec             up;
  MpiBufferPtr            mpibuf[4];

  RefChunk(const nixo_geom* g, int ns, const int* np_required, const double* q, const double* m)
      : Chunk(Dims3D{g->dims[0], g->dims[1], g->dims[2]}, Bool3D{true, true, true}, 0),
        order(g->order), Ns(ns)
  {
    set_boundary_margin(g->nb);
    set_global_context(g->offset, g->gdims);
    set_coordinate(g->del[0], g->del[1], g->del[2]);

    for (int i = 0; i < nbsize; i++) {
      nbid[i]   = 0;
      nbrank[i] = 0;
    }

    size_t mz = dims[0] + 2 * boundary_margin;
    size_t my = dims[1] + 2 * boundary_margin;
    size_t mx = dims[2] + 2 * boundary_margin;
    uf.resize({mz, my, mx, 6ul});
    uj.resize({mz, my, mx, 4ul});
    uf.fill(0);
    uj.fill(0);
    um.resize({mz, my, mx, (size_t)ns, 14ul});
    um.fill(0);

    for (int is = 0; is < ns; is++) {
      auto p = std::make_shared<XtensorParticle>(np_required[is], *this);
      p->q   = q[is];
      p->m   = m[is];
      up.push_back(p);
    }

    // same buffer layout as the reference's tests (test_xtensor_halo3d.cpp:81-82,139-140,189-191)
    mpibuf[NIXO_MODE_FIELD] = std::make_shared<MpiBuffer>();
    set_mpi_buffer(mpibuf[NIXO_MODE_FIELD], 0, 0, sizeof(float64) * 6);
    mpibuf[NIXO_MODE_CURRENT] = std::make_shared<MpiBuffer>();
    set_mpi_buffer(mpibuf[NIXO_MODE_CURRENT], 0, 0, sizeof(float64) * 4);
    mpibuf[NIXO_MODE_PARTICLE] = std::make_shared<MpiBuffer>();
    set_mpi_buffer(mpibuf[NIXO_MODE_PARTICLE], 0, XtensorHaloParticle3D<RefChunk>::head_byte,
                   XtensorHaloParticle3D<RefChunk>::elem_byte);
    mpibuf[NIXO_MODE_MOMENT] = std::make_shared<MpiBuffer>();
    set_mpi_buffer(mpibuf[NIXO_MODE_MOMENT], 0, 0, sizeof(float64) * ns * 14);
  }

  int get_order() const
  {
    return order;
  }

  void setup(json& config) override
  {
  }

  template <int Order, bool Simd>
  void push_deposit(int is, float64 delt, float64 cc);
};

// pusher of the composed step: 0 = push_boris, 1 = push_vay, 2 = push_higuera_cary
// (primitives.hpp:165-253; the reference offers the three as interchangeable primitives)
static int g_pusher = 0;

//
// composed per-particle step, scalar instantiation of the reference templates
//
template <int Order>
static void push_deposit_scalar(RefChunk& c, XtensorParticle& p, int ip, float64 delt, float64 cc)
{
  using namespace nix::primitives;
  constexpr int is_odd = Order % 2;
  constexpr int half   = Order / 2;

  auto& xu = p.xu;
  auto& xv = p.xv;

  const auto [Lbx, Ubx] = c.get_xbound();
  const auto [Lby, Uby] = c.get_ybound();
  const auto [Lbz, Ubz] = c.get_zbound();
  const float64 delx = p.delx, dely = p.dely, delz = p.delz;
  const float64 rdx = 1 / delx, rdy = 1 / dely, rdz = 1 / delz;
  const float64 rc   = 1 / cc;
  const float64 dt1  = 0.5 * p.q / p.m * delt;
  const float64 dxdt = delx / delt, dydt = dely / delt, dzdt = delz / delt;

  // bin offsets exactly as XtensorParticle::count (xtensor_particle.hpp:332-334)
  const float64 xoff  = p.xmin - 0.5 * delx * is_odd;
  const float64 yoff  = p.ymin - 0.5 * dely * is_odd;
  const float64 zoff  = p.zmin - 0.5 * delz * is_odd;
  const float64 xhoff = p.xmin - 0.5 * delx * (1 - is_odd);
  const float64 yhoff = p.ymin - 0.5 * dely * (1 - is_odd);
  const float64 zhoff = p.zmin - 0.5 * delz * (1 - is_odd);
  // position of integer node 0 (cell centre) and of half node 0 (cell edge)
  const float64 ximin = p.xmin + 0.5 * delx;
  const float64 yimin = p.ymin + 0.5 * dely;
  const float64 zimin = p.zmin + 0.5 * delz;

  float64 x = xu(ip, 0), y = xu(ip, 1), z = xu(ip, 2);
  float64 ux = xu(ip, 3), uy = xu(ip, 4), uz = xu(ip, 5);

  // reference nodes of the integer (cell-centre) grid and of the half (cell-edge) grid
  int ix = digitize(x, xoff, rdx) - is_odd;
  int iy = digitize(y, yoff, rdy) - is_odd;
  int iz = digitize(z, zoff, rdz) - is_odd;
  int hx = digitize(x, xhoff, rdx);
  int hy = digitize(y, yhoff, rdy);
  int hz = digitize(z, zhoff, rdz);

  float64 wix[Order + 2] = {0}, wiy[Order + 2] = {0}, wiz[Order + 2] = {0};
  float64 whx[Order + 2] = {0}, why[Order + 2] = {0}, whz[Order + 2] = {0};
  shape_mc<Order>(x, ximin + ix * delx, rdx, wix);
  shape_mc<Order>(y, yimin + iy * dely, rdy, wiy);
  shape_mc<Order>(z, zimin + iz * delz, rdz, wiz);
  shape_mc<Order>(x, p.xmin + hx * delx, rdx, whx);
  shape_mc<Order>(y, p.ymin + hy * dely, rdy, why);
  shape_mc<Order>(z, p.zmin + hz * delz, rdz, whz);

  // common (Order+2)-wide stencil based on the integer grid; half-grid weights shifted into it
  int ix0 = ix - half + Lbx, iy0 = iy - half + Lby, iz0 = iz - half + Lbz;
  interp::shift_weights<Order>(hx - ix, whx);
  interp::shift_weights<Order>(hy - iy, why);
  interp::shift_weights<Order>(hz - iz, whz);

  float64 ex = interp::interp3d<Order>(c.uf, iz0, iy0, ix0, 0, wiz, wiy, whx, dt1);
  float64 ey = interp::interp3d<Order>(c.uf, iz0, iy0, ix0, 1, wiz, why, wix, dt1);
  float64 ez = interp::interp3d<Order>(c.uf, iz0, iy0, ix0, 2, whz, wiy, wix, dt1);
  float64 bx = interp::interp3d<Order>(c.uf, iz0, iy0, ix0, 3, whz, why, wix, dt1);
  float64 by = interp::interp3d<Order>(c.uf, iz0, iy0, ix0, 4, whz, wiy, whx, dt1);
  float64 bz = interp::interp3d<Order>(c.uf, iz0, iy0, ix0, 5, wiz, why, whx, dt1);

  if (g_pusher == 1) push_vay(ux, uy, uz, ex, ey, ez, bx, by, bz, cc);
  else if (g_pusher == 2) push_higuera_cary(ux, uy, uz, ex, ey, ez, bx, by, bz, cc);
  else push_boris(ux, uy, uz, ex, ey, ez, bx, by, bz, cc);

  float64 gam = lorentz_factor(ux, uy, uz, rc);
  float64 dtg = delt / gam;

  // keep the old position in xv, advance xu (test_esirkepov.cpp:1046-1051)
  xv(ip, 0) = x;
  xv(ip, 1) = y;
  xv(ip, 2) = z;
  xu(ip, 0) = x + ux * dtg;
  xu(ip, 1) = y + uy * dtg;
  xu(ip, 2) = z + uz * dtg;
  xu(ip, 3) = ux;
  xu(ip, 4) = uy;
  xu(ip, 5) = uz;

  // Esirkepov deposit (test_esirkepov.cpp:1056-1088)
  float64 ss[2][3][Order + 3] = {0};
  shape_mc<Order>(xv(ip, 0), ximin + ix * delx, rdx, &ss[0][0][1]);
  shape_mc<Order>(xv(ip, 1), yimin + iy * dely, rdy, &ss[0][1][1]);
  shape_mc<Order>(xv(ip, 2), zimin + iz * delz, rdz, &ss[0][2][1]);

  int ix1 = digitize(xu(ip, 0), xoff, rdx) - is_odd;
  int iy1 = digitize(xu(ip, 1), yoff, rdy) - is_odd;
  int iz1 = digitize(xu(ip, 2), zoff, rdz) - is_odd;
  if (std::abs(ix1 - ix) > 1 || std::abs(iy1 - iy) > 1 || std::abs(iz1 - iz) > 1) {
    return; // c*dt > dx : outside the scheme's validity (nothing deposited)
  }
  shape_mc<Order>(xu(ip, 0), ximin + ix1 * delx, rdx, &ss[1][0][1 + ix1 - ix]);
  shape_mc<Order>(xu(ip, 1), yimin + iy1 * dely, rdy, &ss[1][1][1 + iy1 - iy]);
  shape_mc<Order>(xu(ip, 2), zimin + iz1 * delz, rdz, &ss[1][2][1 + iz1 - iz]);

  float64 cur[Order + 3][Order + 3][Order + 3][4] = {0};
  esirkepov::deposit3d<Order>(dxdt, dydt, dzdt, p.q, ss, cur);

  int jx0 = ix - half - 1 + Lbx, jy0 = iy - half - 1 + Lby, jz0 = iz - half - 1 + Lbz;
  append_current3d<Order>(c.uj, jz0, jy0, jx0, cur);
}

//
// the reference's vectorised *sorted* code paths: one xsimd batch = consecutive particles of ONE
// cell (scalar stencil index -> interp3d_impl_sorted, append_current3d reduce_add branch).
// Requires a cell-sorted container (pindex valid).  Remainders use the scalar instantiation.
//
template <int Order>
static void push_deposit_simd_cell(RefChunk& c, XtensorParticle& p, int ip0, float64 delt,
                                   float64 cc)
{
  using namespace nix::primitives;
  using V              = simd_f64;
  using I              = simd_i64;
  constexpr int is_odd = Order % 2;
  constexpr int half   = Order / 2;
  constexpr int W      = V::size;

  auto& xu = p.xu;
  auto& xv = p.xv;

  const auto [Lbx, Ubx] = c.get_xbound();
  const auto [Lby, Uby] = c.get_ybound();
  const auto [Lbz, Ubz] = c.get_zbound();
  const float64 delx = p.delx, dely = p.dely, delz = p.delz;
  const V       rdx = 1 / delx, rdy = 1 / dely, rdz = 1 / delz;
  const V       rc   = 1 / cc;
  const V       dt1  = 0.5 * p.q / p.m * delt;
  const float64 dxdt = delx / delt, dydt = dely / delt, dzdt = delz / delt;
  const V       xoff  = p.xmin - 0.5 * delx * is_odd;
  const V       yoff  = p.ymin - 0.5 * dely * is_odd;
  const V       zoff  = p.zmin - 0.5 * delz * is_odd;
  const V       xhoff = p.xmin - 0.5 * delx * (1 - is_odd);
  const V       yhoff = p.ymin - 0.5 * dely * (1 - is_odd);
  const V       zhoff = p.zmin - 0.5 * delz * (1 - is_odd);
  const float64 ximin = p.xmin + 0.5 * delx;
  const float64 yimin = p.ymin + 0.5 * dely;
  const float64 zimin = p.zmin + 0.5 * delz;

  const I index = xsimd::detail::make_sequence_as_batch<I>() * 7;
  V       q[6];
  for (int k = 0; k < 6; k++) {
    q[k] = V::gather(&xu(ip0, k), index);
  }
  V x = q[0], y = q[1], z = q[2], ux = q[3], uy = q[4], uz = q[5];

  auto ixv = digitize(x, xoff, rdx) - is_odd;
  auto iyv = digitize(y, yoff, rdy) - is_odd;
  auto izv = digitize(z, zoff, rdz) - is_odd;
  auto hxv = digitize(x, xhoff, rdx);
  auto hyv = digitize(y, yhoff, rdy);
  auto hzv = digitize(z, zhoff, rdz);

  // all lanes share the stencil of the cell (sorted container)
  int ix = ixv.get(0), iy = iyv.get(0), iz = izv.get(0);

  V wix[Order + 2] = {0}, wiy[Order + 2] = {0}, wiz[Order + 2] = {0};
  V whx[Order + 2] = {0}, why[Order + 2] = {0}, whz[Order + 2] = {0};
  shape_mc<Order>(x, xsimd::to_float(ixv) * delx + ximin, rdx, wix);
  shape_mc<Order>(y, xsimd::to_float(iyv) * dely + yimin, rdy, wiy);
  shape_mc<Order>(z, xsimd::to_float(izv) * delz + zimin, rdz, wiz);
  shape_mc<Order>(x, xsimd::to_float(hxv) * delx + p.xmin, rdx, whx);
  shape_mc<Order>(y, xsimd::to_float(hyv) * dely + p.ymin, rdy, why);
  shape_mc<Order>(z, xsimd::to_float(hzv) * delz + p.zmin, rdz, whz);

  int ix0 = ix - half + Lbx, iy0 = iy - half + Lby, iz0 = iz - half + Lbz;
  interp::shift_weights<Order>(hxv - ixv, whx);
  interp::shift_weights<Order>(hyv - iyv, why);
  interp::shift_weights<Order>(hzv - izv, whz);

  V ex = interp::interp3d<Order>(c.uf, iz0, iy0, ix0, 0, wiz, wiy, whx, dt1);
  V ey = interp::interp3d<Order>(c.uf, iz0, iy0, ix0, 1, wiz, why, wix, dt1);
  V ez = interp::interp3d<Order>(c.uf, iz0, iy0, ix0, 2, whz, wiy, wix, dt1);
  V bx = interp::interp3d<Order>(c.uf, iz0, iy0, ix0, 3, whz, why, wix, dt1);
  V by = interp::interp3d<Order>(c.uf, iz0, iy0, ix0, 4, whz, wiy, whx, dt1);
  V bz = interp::interp3d<Order>(c.uf, iz0, iy0, ix0, 5, wiz, why, whx, dt1);

  if (g_pusher == 1) push_vay(ux, uy, uz, ex, ey, ez, bx, by, bz, V(cc));
  else if (g_pusher == 2) push_higuera_cary(ux, uy, uz, ex, ey, ez, bx, by, bz, V(cc));
  else push_boris(ux, uy, uz, ex, ey, ez, bx, by, bz, V(cc));

  V gam = lorentz_factor(ux, uy, uz, rc);
  V dtg = V(delt) / gam;
  V xn = x + ux * dtg, yn = y + uy * dtg, zn = z + uz * dtg;

  x.scatter(&xv(ip0, 0), index);
  y.scatter(&xv(ip0, 1), index);
  z.scatter(&xv(ip0, 2), index);
  xn.scatter(&xu(ip0, 0), index);
  yn.scatter(&xu(ip0, 1), index);
  zn.scatter(&xu(ip0, 2), index);
  ux.scatter(&xu(ip0, 3), index);
  uy.scatter(&xu(ip0, 4), index);
  uz.scatter(&xu(ip0, 5), index);

  V ss[2][3][Order + 3] = {0};
  shape_mc<Order>(x, xsimd::to_float(ixv) * delx + ximin, rdx, &ss[0][0][1]);
  shape_mc<Order>(y, xsimd::to_float(iyv) * dely + yimin, rdy, &ss[0][1][1]);
  shape_mc<Order>(z, xsimd::to_float(izv) * delz + zimin, rdz, &ss[0][2][1]);

  auto ix1 = digitize(xn, xoff, rdx) - is_odd;
  auto iy1 = digitize(yn, yoff, rdy) - is_odd;
  auto iz1 = digitize(zn, zoff, rdz) - is_odd;
  shape_mc<Order>(xn, xsimd::to_float(ix1) * delx + ximin, rdx, &ss[1][0][1]);
  shape_mc<Order>(yn, xsimd::to_float(iy1) * dely + yimin, rdy, &ss[1][1][1]);
  shape_mc<Order>(zn, xsimd::to_float(iz1) * delz + zimin, rdz, &ss[1][2][1]);

  // in-place shift of ss[1] according to particle movement (test_esirkepov.cpp:1169-1170)
  xsimd::batch<int64_t> shift[3] = {ix1 - ixv, iy1 - iyv, iz1 - izv};
  esirkepov::shift_weights<3, Order>(shift, ss[1]);

  V cur[Order + 3][Order + 3][Order + 3][4] = {0};
  esirkepov::deposit3d<Order>(dxdt, dydt, dzdt, V(p.q), ss, cur);

  int jx0 = ix - half - 1 + Lbx, jy0 = iy - half - 1 + Lby, jz0 = iz - half - 1 + Lbz;
  append_current3d<Order>(c.uj, jz0, jy0, jx0, cur);
  (void)W;
}

template <int Order, bool Simd>
void RefChunk::push_deposit(int is, float64 delt, float64 cc)
{
  auto& p = *up[is];

  if constexpr (Simd == false) {
    for (int ip = 0; ip < p.Np; ip++) {
      push_deposit_scalar<Order>(*this, p, ip, delt, cc);
    }
  } else {
    // cell by cell over the sorted container (pindex from XtensorParticle::sort)
    constexpr int W = simd_f64::size;
    const int     nx = Ubx - Lbx + 2, ny = Uby - Lby + 2, nz = Ubz - Lbz + 2;
    const int     ncell = nx * ny * nz;
    for (int ii = 0; ii < ncell; ii++) {
      int ip_zero = p.pindex(ii);
      int np_cell = p.pindex(ii + 1) - ip_zero;
      int np_simd = (np_cell / W) * W;
      for (int ip = ip_zero; ip < ip_zero + np_simd; ip += W) {
        push_deposit_simd_cell<Order>(*this, p, ip, delt, cc);
      }
      for (int ip = ip_zero + np_simd; ip < ip_zero + np_cell; ip++) {
        push_deposit_scalar<Order>(*this, p, ip, delt, cc);
      }
    }
  }
}

inline RefChunk* R(nixo_chunk* c)
{
  return reinterpret_cast<RefChunk*>(c);
}
} // namespace

namespace
{
// per-particle composition documented in nix_oracle.c (the reference ships the scatter, not the loop)
template <int Order>
void deposit_moment_t(RefChunk& c, float64 cc)
{
  using namespace nix::primitives;
  constexpr int is_odd = Order % 2, half = Order / 2, n1 = Order + 1;
  const auto [Lbx, Ubx] = c.get_xbound();
  const auto [Lby, Uby] = c.get_ybound();
  const auto [Lbz, Ubz] = c.get_zbound();
  const float64 rc = 1 / cc;
  c.um.fill(0);
  for (int is = 0; is < c.Ns; is++) {
    auto&         p    = *c.up[is];
    const float64 del[3] = {p.delz, p.dely, p.delx};
    const float64 lo[3]  = {p.zmin, p.ymin, p.xmin};
    const int     Lb[3]  = {Lbz, Lby, Lbx};
    for (int ip = 0; ip < p.Np; ip++) {
      const float64 pos[3] = {p.xu(ip, 2), p.xu(ip, 1), p.xu(ip, 0)};
      float64       w[3][n1];
      int           i0[3];
      for (int a = 0; a < 3; a++) {
        const float64 rdx = 1 / del[a];
        int           i   = digitize(pos[a], lo[a] - 0.5 * del[a] * is_odd, rdx) - is_odd;
        shape_mc<Order>(pos[a], (lo[a] + 0.5 * del[a]) + i * del[a], rdx, w[a]);
        i0[a] = i - half + Lb[a];
      }
      const float64 ux = p.xu(ip, 3), uy = p.xu(ip, 4), uz = p.xu(ip, 5), m = p.m;
      const float64 gam = lorentz_factor(ux, uy, uz, rc);
      const float64 mom[14] = {m,           m * ux / gam, m * uy / gam, m * uz / gam, m * gam * cc * cc,
                               m * ux * cc, m * uy * cc,  m * uz * cc,  m * ux * ux / gam, m * uy * uy / gam,
                               m * uz * uz / gam, m * ux * uy / gam, m * uy * uz / gam, m * uz * ux / gam};
      float64 moment[n1][n1][n1][14];
      for (int jz = 0; jz < n1; jz++)
        for (int jy = 0; jy < n1; jy++)
          for (int jx = 0; jx < n1; jx++) {
            const float64 ww = (w[0][jz] * w[1][jy]) * w[2][jx];
            for (int k = 0; k < 14; k++) moment[jz][jy][jx][k] = ww * mom[k];
          }
      append_moment3d<Order>(c.um, i0[0], i0[1], i0[2], is, moment);
    }
  }
}

struct PackData {
  int Lbz, Ubz, Lby, Uby, Lbx, Ubx;
};
PackData pack_data(RefChunk& c)
{
  const auto [Lbx, Ubx] = c.get_xbound();
  const auto [Lby, Uby] = c.get_ybound();
  const auto [Lbz, Ubz] = c.get_zbound();
  return PackData{Lbz, Ubz, Lby, Uby, Lbx, Ubx};
}
} // namespace

template <int Order>
static void ref_append_current3d(double* uj, int my, int mx, int iz0, int iy0, int ix0, const double* cur)
{
  constexpr int n  = Order + 3;
  size_t        mz = static_cast<size_t>(iz0 + n);
  auto view = xt::adapt(uj, mz * my * mx * 4, xt::no_ownership(), std::vector<size_t>{mz, (size_t)my, (size_t)mx, 4ul});
  double local[n][n][n][4];
  std::memcpy(local, cur, sizeof(local));
  primitives::append_current3d<Order>(view, iz0, iy0, ix0, local);
}

template <int Order>
static void ref_append_moment3d(double* um, int my, int mx, int ns, int iz0, int iy0, int ix0, int is, const double* mom)
{
  constexpr int n  = Order + 1;
  size_t        mz = static_cast<size_t>(iz0 + n);
  auto view = xt::adapt(um, mz * my * mx * ns * 14, xt::no_ownership(),
                        std::vector<size_t>{mz, (size_t)my, (size_t)mx, (size_t)ns, 14ul});
  double local[n][n][n][14];
  std::memcpy(local, mom, sizeof(local));
  primitives::append_moment3d<Order>(view, iz0, iy0, ix0, is, local);
}

extern "C" {

const char* nixo_impl_name(void)
{
  return "reference";
}

int nixo_simd_lanes(void)
{
  return simd_f64::size;
}

nixo_chunk* nixo_chunk_create(const nixo_geom* g, int ns, const int* np_required, const double* q,
                              const double* m)
{
  return reinterpret_cast<nixo_chunk*>(new RefChunk(g, ns, np_required, q, m));
}

void nixo_chunk_destroy(nixo_chunk* c)
{
  delete R(c);
}

double* nixo_chunk_uf(nixo_chunk* c)
{
  return R(c)->uf.data();
}

double* nixo_chunk_uj(nixo_chunk* c)
{
  return R(c)->uj.data();
}

void nixo_chunk_set_nb_valid(nixo_chunk* c, int iz, int iy, int ix, int valid)
{
  R(c)->set_nb_rank(iz - 1, iy - 1, ix - 1, valid ? 0 : MPI_PROC_NULL);
}

int nixo_particle_ng(nixo_chunk* c, int is)
{
  return R(c)->up[is]->Ng;
}
int nixo_particle_np(nixo_chunk* c, int is)
{
  return R(c)->up[is]->Np;
}
void nixo_particle_set_np(nixo_chunk* c, int is, int np)
{
  R(c)->up[is]->Np = np;
}
int nixo_particle_np_total(nixo_chunk* c, int is)
{
  return R(c)->up[is]->Np_total;
}
double* nixo_particle_xu(nixo_chunk* c, int is)
{
  return R(c)->up[is]->xu.data();
}
double* nixo_particle_xv(nixo_chunk* c, int is)
{
  return R(c)->up[is]->xv.data();
}
int32_t* nixo_particle_gindex(nixo_chunk* c, int is)
{
  return R(c)->up[is]->gindex.data();
}
int32_t* nixo_particle_pindex(nixo_chunk* c, int is)
{
  return R(c)->up[is]->pindex.data();
}
int32_t* nixo_particle_pcount(nixo_chunk* c, int is)
{
  return R(c)->up[is]->pcount.data();
}
void nixo_particle_resize(nixo_chunk* c, int is, int np_required)
{
  R(c)->up[is]->resize(np_required);
}
void nixo_particle_count(nixo_chunk* c, int is, int lbp, int ubp, int reset, int order)
{
  R(c)->up[is]->count(lbp, ubp, reset != 0, order);
}
void nixo_particle_sort(nixo_chunk* c, int is)
{
  R(c)->up[is]->sort();
}
void nixo_particle_set_boundary_periodic(nixo_chunk* c, int is, int lbp, int ubp)
{
  R(c)->up[is]->set_boundary_periodic(lbp, ubp);
}

int nixo_digitize(double x, double xmin, double rdx)
{
  return primitives::digitize(x, xmin, rdx);
}

void nixo_shape_mc(int order, double x, double X, double rdx, double* s)
{
  switch (order) {
  case 1:
    primitives::shape_mc<1>(x, X, rdx, s);
    break;
  case 2:
    primitives::shape_mc<2>(x, X, rdx, s);
    break;
  case 3:
    primitives::shape_mc<3>(x, X, rdx, s);
    break;
  case 4:
    primitives::shape_mc<4>(x, X, rdx, s);
    break;
  default:
    break;
  }
}

void nixo_shape_wt(int order, double x, double X, double rdx, double dt, double rdt, double* s)
{
  switch (order) {
  case 1:
    primitives::shape_wt<1>(x, X, rdx, dt, rdt, s);
    break;
  case 2:
    primitives::shape_wt<2>(x, X, rdx, dt, rdt, s);
    break;
  case 3:
    primitives::shape_wt<3>(x, X, rdx, dt, rdt, s);
    break;
  case 4:
    primitives::shape_wt<4>(x, X, rdx, dt, rdt, s);
    break;
  default:
    break;
  }
}

void nixo_push_boris(double* u, const double* eb, double cc)
{
  primitives::push_boris(u[0], u[1], u[2], eb[0], eb[1], eb[2], eb[3], eb[4], eb[5], cc);
}
void nixo_push_vay(double* u, const double* eb, double cc)
{
  primitives::push_vay(u[0], u[1], u[2], eb[0], eb[1], eb[2], eb[3], eb[4], eb[5], cc);
}
void nixo_push_higuera_cary(double* u, const double* eb, double cc)
{
  primitives::push_higuera_cary(u[0], u[1], u[2], eb[0], eb[1], eb[2], eb[3], eb[4], eb[5], cc);
}
void nixo_set_pusher(int pusher) { g_pusher = pusher; }
int  nixo_get_pusher(void) { return g_pusher; }

double nixo_lorentz_factor(double ux, double uy, double uz, double rc)
{
  return primitives::lorentz_factor(ux, uy, uz, rc);
}

void nixo_deposit3d(int order, double dxdt, double dydt, double dzdt, double qs, double* ss,
                    double* cur)
{
  switch (order) {
  case 1:
    esirkepov::deposit3d<1>(dxdt, dydt, dzdt, qs, reinterpret_cast<double(*)[3][4]>(ss),
                            reinterpret_cast<double(*)[4][4][4]>(cur));
    break;
  case 2:
    esirkepov::deposit3d<2>(dxdt, dydt, dzdt, qs, reinterpret_cast<double(*)[3][5]>(ss),
                            reinterpret_cast<double(*)[5][5][4]>(cur));
    break;
  case 3:
    esirkepov::deposit3d<3>(dxdt, dydt, dzdt, qs, reinterpret_cast<double(*)[3][6]>(ss),
                            reinterpret_cast<double(*)[6][6][4]>(cur));
    break;
  case 4:
    esirkepov::deposit3d<4>(dxdt, dydt, dzdt, qs, reinterpret_cast<double(*)[3][7]>(ss),
                            reinterpret_cast<double(*)[7][7][4]>(cur));
    break;
  default:
    break;
  }
}

void nixo_append_current3d(int order, double* uj, int my, int mx, int iz0, int iy0, int ix0, const double* cur)
{
  switch (order) {
  case 1: ref_append_current3d<1>(uj, my, mx, iz0, iy0, ix0, cur); break;
  case 2: ref_append_current3d<2>(uj, my, mx, iz0, iy0, ix0, cur); break;
  case 3: ref_append_current3d<3>(uj, my, mx, iz0, iy0, ix0, cur); break;
  case 4: ref_append_current3d<4>(uj, my, mx, iz0, iy0, ix0, cur); break;
  default: break;
  }
}

void nixo_append_moment3d(int order, double* um, int my, int mx, int ns, int iz0, int iy0, int ix0, int is,
                          const double* mom)
{
  switch (order) {
  case 1: ref_append_moment3d<1>(um, my, mx, ns, iz0, iy0, ix0, is, mom); break;
  case 2: ref_append_moment3d<2>(um, my, mx, ns, iz0, iy0, ix0, is, mom); break;
  case 3: ref_append_moment3d<3>(um, my, mx, ns, iz0, iy0, ix0, is, mom); break;
  case 4: ref_append_moment3d<4>(um, my, mx, ns, iz0, iy0, ix0, is, mom); break;
  default: break;
  }
}

void nixo_interp_shift_weights(int order, int shift, double* ww)
{
  switch (order) {
  case 1:
    interp::shift_weights<1>(shift, ww);
    break;
  case 2:
    interp::shift_weights<2>(shift, ww);
    break;
  case 3:
    interp::shift_weights<3>(shift, ww);
    break;
  case 4:
    interp::shift_weights<4>(shift, ww);
    break;
  default:
    break;
  }
}

void nixo_esirkepov_shift_weights(int order, const int* shift, double* ss)
{
  int sh[3] = {shift[0], shift[1], shift[2]};
  switch (order) {
  case 1:
    esirkepov::shift_weights<3, 1>(sh, reinterpret_cast<double(*)[4]>(ss));
    break;
  case 2:
    esirkepov::shift_weights<3, 2>(sh, reinterpret_cast<double(*)[5]>(ss));
    break;
  case 3:
    esirkepov::shift_weights<3, 3>(sh, reinterpret_cast<double(*)[6]>(ss));
    break;
  case 4:
    esirkepov::shift_weights<3, 4>(sh, reinterpret_cast<double(*)[7]>(ss));
    break;
  default:
    break;
  }
}

double nixo_interp3d(int order, const double* eb, int my, int mx, int iz0, int iy0, int ix0, int ik,
                     const double* wz, const double* wy, const double* wx, double dt)
{
  // wrap the raw array as an xtensor view of shape [*][my][mx][6]
  size_t mz   = static_cast<size_t>(iz0 + order + 2);
  auto   view = xt::adapt(eb, mz * my * mx * 6, xt::no_ownership(),
                          std::vector<size_t>{mz, (size_t)my, (size_t)mx, 6ul});
  double* z = const_cast<double*>(wz);
  double* y = const_cast<double*>(wy);
  double* x = const_cast<double*>(wx);
  switch (order) {
  case 1:
    return interp::interp3d<1>(view, iz0, iy0, ix0, ik, z, y, x, dt);
  case 2:
    return interp::interp3d<2>(view, iz0, iy0, ix0, ik, z, y, x, dt);
  case 3:
    return interp::interp3d<3>(view, iz0, iy0, ix0, ik, z, y, x, dt);
  case 4:
    return interp::interp3d<4>(view, iz0, iy0, ix0, ik, z, y, x, dt);
  default:
    return 0;
  }
}

void nixo_chunk_push_deposit(nixo_chunk* c, double delt, double cc, int simd)
{
  RefChunk* r = R(c);
  for (int is = 0; is < r->Ns; is++) {
    if (simd) {
      switch (r->order) {
      case 1:
        r->push_deposit<1, true>(is, delt, cc);
        break;
      case 2:
        r->push_deposit<2, true>(is, delt, cc);
        break;
      case 3:
        r->push_deposit<3, true>(is, delt, cc);
        break;
      case 4:
        r->push_deposit<4, true>(is, delt, cc);
        break;
      }
    } else {
      switch (r->order) {
      case 1:
        r->push_deposit<1, false>(is, delt, cc);
        break;
      case 2:
        r->push_deposit<2, false>(is, delt, cc);
        break;
      case 3:
        r->push_deposit<3, false>(is, delt, cc);
        break;
      case 4:
        r->push_deposit<4, false>(is, delt, cc);
        break;
      }
    }
  }
}

void nixo_chunk_halo_pack(nixo_chunk* c, int mode)
{
  RefChunk* r = R(c);
  if (mode == NIXO_MODE_FIELD) {
    XtensorHaloField3D<RefChunk> halo(r->uf, *r);
    r->pack_bc_exchange(r->mpibuf[mode], halo);
  } else if (mode == NIXO_MODE_CURRENT) {
    XtensorHaloCurrent3D<RefChunk> halo(r->uj, *r);
    r->pack_bc_exchange(r->mpibuf[mode], halo);
  } else if (mode == NIXO_MODE_MOMENT) {
    XtensorHaloMoment3D<RefChunk> halo(r->um, *r);
    r->pack_bc_exchange(r->mpibuf[mode], halo);
  } else {
    XtensorHaloParticle3D<RefChunk> halo(r->up, *r);
    r->pack_bc_exchange(r->mpibuf[mode], halo);
  }
}

void nixo_chunk_halo_unpack(nixo_chunk* c, int mode)
{
  RefChunk* r = R(c);
  if (mode == NIXO_MODE_FIELD) {
    XtensorHaloField3D<RefChunk> halo(r->uf, *r);
    r->unpack_bc_exchange(r->mpibuf[mode], halo);
  } else if (mode == NIXO_MODE_CURRENT) {
    XtensorHaloCurrent3D<RefChunk> halo(r->uj, *r);
    r->unpack_bc_exchange(r->mpibuf[mode], halo);
  } else if (mode == NIXO_MODE_MOMENT) {
    XtensorHaloMoment3D<RefChunk> halo(r->um, *r);
    r->unpack_bc_exchange(r->mpibuf[mode], halo);
  } else {
    XtensorHaloParticle3D<RefChunk> halo(r->up, *r);
    r->unpack_bc_exchange(r->mpibuf[mode], halo);
  }
}

int nixo_chunk_bufsize(nixo_chunk* c, int mode, int iz, int iy, int ix)
{
  return R(c)->mpibuf[mode]->bufsize(iz, iy, ix);
}

int nixo_chunk_bufaddr(nixo_chunk* c, int mode, int iz, int iy, int ix)
{
  return R(c)->mpibuf[mode]->bufaddr(iz, iy, ix);
}

uint8_t* nixo_chunk_sendbuf(nixo_chunk* c, int mode)
{
  return R(c)->mpibuf[mode]->sendbuf.get(0);
}

int nixo_chunk_sendbuf_size(nixo_chunk* c, int mode)
{
  return R(c)->mpibuf[mode]->sendbuf.size;
}

void nixo_chunk_set_recv_sizes(nixo_chunk* c, int mode, const int* bufsize27)
{
  // what probe_bc_exchange does once every message has been probed (chunk.cpp:356-368)
  auto mpibuf  = R(c)->mpibuf[mode];
  int  bufsize = 0;
  for (int iz = 0; iz <= 2; iz++) {
    for (int iy = 0; iy <= 2; iy++) {
      for (int ix = 0; ix <= 2; ix++) {
        mpibuf->bufsize(iz, iy, ix) = bufsize27[9 * iz + 3 * iy + ix];
        mpibuf->bufaddr(iz, iy, ix) = bufsize;
        bufsize += mpibuf->bufsize(iz, iy, ix);
      }
    }
  }
  mpibuf->recvbuf.resize(bufsize);
}

uint8_t* nixo_chunk_recvbuf(nixo_chunk* c, int mode)
{
  return R(c)->mpibuf[mode]->recvbuf.get(0);
}

int nixo_chunk_recvbuf_size(nixo_chunk* c, int mode)
{
  return R(c)->mpibuf[mode]->recvbuf.size;
}

// ---- moments and diagnostic packers: the reference's own append_moment3d / XtensorPacker3D ----
double* nixo_chunk_um(nixo_chunk* c)
{
  return R(c)->um.data();
}

void nixo_chunk_deposit_moment(nixo_chunk* c, double cc)
{
  RefChunk* r = R(c);
  switch (r->order) {
  case 1: deposit_moment_t<1>(*r, cc); break;
  case 2: deposit_moment_t<2>(*r, cc); break;
  case 3: deposit_moment_t<3>(*r, cc); break;
  case 4: deposit_moment_t<4>(*r, cc); break;
  }
}

int nixo_chunk_pack_field(nixo_chunk* c, int decimate, double* out)
{
  RefChunk*       r = R(c);
  XtensorPacker3D packer;
  return (int)(packer.pack_field(r->uf, pack_data(*r), decimate, reinterpret_cast<uint8_t*>(out), 0) / sizeof(float64));
}

int nixo_chunk_pack_moment(nixo_chunk* c, int which, int decimate, double* out)
{
  RefChunk*       r = R(c);
  XtensorPacker3D packer;
  size_t          n = which == 0 ? packer.pack_moment(r->uj, pack_data(*r), decimate, reinterpret_cast<uint8_t*>(out), 0)
                                 : packer.pack_moment(r->um, pack_data(*r), decimate, reinterpret_cast<uint8_t*>(out), 0);
  return (int)(n / sizeof(float64));
}

int nixo_chunk_pack_tracer(nixo_chunk* c, int is, double* out)
{
  RefChunk*       r = R(c);
  XtensorPacker3D packer;
  return (int)(packer.pack_tracer(r->up[is], reinterpret_cast<uint8_t*>(out), 0) / (sizeof(float64) * 7));
}

} // extern "C"
