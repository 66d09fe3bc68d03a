/* nix_oracle.c -- plain-C restatement of the per-chunk PIC step of amanotk/nix.
 *
 * TEST INFRASTRUCTURE ONLY (see nix_oracle.h).  Every function cites the reference lines it
 * restates (paths relative to the reference root).  Compiled with -ffp-contract=off so that every
 * floating-point expression is evaluated exactly as written; the expressions keep the reference's
 * association order, and tests/test_oracle_vs_ref.py pins this file BIT FOR BIT against the
 * reference's own templates (oracle/_ref) and against the golden vectors in tests/golden/.
 *
 * Parity status: PINNED (golden vectors generated from the reference itself by
 * tests/golden/make_golden.py + the reference's known-answer tests restated in tests/).
 */
#include "nix_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NC 7          /* particle.hpp:18  */
#define ALLOC_UNIT 128 /* particle.hpp:19  */
#define LANES 8       /* nix.hpp:105-108  */
#define MAXO 4
#define HEAD_BYTE 4   /* xtensor_halo3d.hpp:258 */
#define ELEM_BYTE 56  /* xtensor_halo3d.hpp:259 */

typedef struct {
  int      Np_total, Np, Ng;
  double   q, m;
  double*  xu;
  double*  xv;
  int32_t* gindex;
  int32_t* pindex;
  int32_t* pcount;
} particle_t;

typedef struct {
  int      bufsize[27];
  int      bufaddr[27];
  uint8_t* sendbuf;
  int      sendsize;
  uint8_t* recvbuf;
  int      recvsize;
} mpibuf_t;

struct nixo_chunk {
  nixo_geom g;
  int       ns;
  int       Lb[3], Ub[3];                 /* z, y, x  (chunk.cpp:134-139) */
  int       sendlb[3][3], sendub[3][3];   /* [axis][dir] (chunk.cpp:171-188) */
  int       recvlb[3][3], recvub[3][3];   /* (chunk.cpp:190-207) */
  double    lim[3][2];                    /* chunk range (chunk.cpp:210-237) */
  double    glim[3][2];                   /* global range */
  int       M[3];
  int       nbvalid[27];
  double*   uf;
  double*   uj;
  double*   um;                           /* [Mz][My][Mx][ns][14] moments */
  particle_t* up;
  mpibuf_t  mpibuf[4];
  int*      num_unpacked;
};

const char* nixo_impl_name(void)
{
  return "port";
}

int nixo_simd_lanes(void)
{
  return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* numerical primitives                                                                       */
/* ------------------------------------------------------------------------------------------ */

/* primitives.hpp:46-58 */
int nixo_digitize(double x, double xmin, double rdx)
{
  return (int)floor((x - xmin) * rdx);
}

/* primitives.hpp:257-298 (shape_mc1/2/3), dispatch :519-532 */
void nixo_shape_mc(int order, double x, double X, double rdx, double* s)
{
  double delta = (x - X) * rdx;
  if (order == 1) {
    s[0] = 1 - delta;
    s[1] = delta;
  } else if (order == 2) {
    double w0 = delta;
    double w1 = 0.5 - w0;
    double w2 = 0.5 + w0;
    s[0]      = 0.50 * w1 * w1;
    s[1]      = 0.75 - w0 * w0;
    s[2]      = 0.50 * w2 * w2;
  } else if (order == 3) {
    const double a       = 1 / 6.0;
    double       w1      = delta;
    double       w2      = 1 - delta;
    double       w1_pow2 = w1 * w1;
    double       w2_pow2 = w2 * w2;
    double       w1_pow3 = w1_pow2 * w1;
    double       w2_pow3 = w2_pow2 * w2;
    s[0]                 = a * w2_pow3;
    s[1]                 = a * (4 - 6 * w1_pow2 + 3 * w1_pow3);
    s[2]                 = a * (4 - 6 * w2_pow2 + 3 * w2_pow3);
    s[3]                 = a * w1_pow3;
  } else if (order == 4) { /* shape_mc4, primitives.hpp:302-331 */
    const double a = 1 / 384.0, b = 1 / 96.0, cc = 115 / 192.0, d = 1 / 8.0;
    double p1 = 1 + delta, m1 = 1 - delta, p2 = 1 + delta * 2, m2 = 1 - delta * 2;
    double d2 = delta * delta;
    double p1_2 = p1 * p1, m1_2 = m1 * m1, p1_3 = p1_2 * p1, m1_3 = m1_2 * m1, p1_4 = p1_3 * p1, m1_4 = m1_3 * m1;
    s[0] = a * (m2 * m2 * m2 * m2);
    s[1] = b * (55 + 20 * p1 - 120 * p1_2 + 80 * p1_3 - 16 * p1_4);
    s[2] = cc + d * d2 * (2 * d2 - 5);
    s[3] = b * (55 + 20 * m1 - 120 * m1_2 + 80 * m1_3 - 16 * m1_4);
    s[4] = a * (p2 * p2 * p2 * p2);
  }
}

/* shape functions of the WT scheme (time-step dependent assignment weights; Lu et al., JCP 413, 109388):
 * primitives.hpp:333-495, dispatch :555-570.  Three branches selected by where delta sits relative to
 * +-dt (even orders) or 1/2 +- dt (odd orders), blended by 0/1 masks exactly as the reference does. */
void nixo_shape_wt(int order, double x, double X, double rdx, double dt, double rdt, double* s)
{
  double delta = (x - X) * rdx;
  if (order == 1) {
    double v = 0.25 * rdt * (1 + 2 * dt - 2 * delta);
    double ss = fmin(1.0, fmax(0.0, v));
    s[0] = ss;
    s[1] = 1 - ss;
    return;
  }
  const int    odd = order & 1;
  const double lo = odd ? 0.5 - dt : -dt, hi = odd ? 0.5 + dt : +dt;
  const double t1 = (delta < lo) ? 1.0 : 0.0, t2 = 1 - t1, t3 = (delta < hi) ? 1.0 : 0.0, t4 = 1 - t3;
  double       A[5] = {0}, B[5] = {0}, C[5] = {0}; /* the three branches */
  if (order == 2) {
    double w0 = fabs(delta), w1 = dt - delta, w2 = dt + delta;
    A[0] = w0, A[1] = 1 - w0, A[2] = 0;
    B[0] = 0.25 * rdt * w1 * w1;
    B[1] = 0.50 * rdt * (dt * (2 - dt) - w0 * w0);
    B[2] = 0.25 * rdt * w2 * w2;
    C[0] = A[2], C[1] = A[1], C[2] = A[0];
  } else if (order == 3) {
    const double a = 1 / 96.0, b = 1 / 24.0, c = 1 / 12.0, adt = a * rdt;
    double w0 = delta, w1 = 1 - delta, w3 = 1 - 2 * delta, w4 = 1 + 2 * delta;
    double w5 = 2 * dt + w3, w6 = 2 * dt - w3, w7 = 3 - 2 * delta;
    double w0_2 = w0 * w0, w1_2 = w1 * w1, w3_2 = w3 * w3, w3_3 = w3_2 * w3, w4_2 = w4 * w4;
    double w5_3 = w5 * w5 * w5, w6_3 = w6 * w6 * w6, w7_2 = w7 * w7;
    double dt2 = dt * dt, dt3 = dt2 * dt, dt2_4 = 4 * dt2;
    double so = adt * (-8 * dt3 - 6 * dt * w3_2), se = adt * (-36 * dt2 * w3 - 3 * w3_3);
    A[0] = b * (dt2_4 + 3 * w3_2), A[1] = c * (9 - dt2_4 - 12 * w0_2), A[2] = b * (dt2_4 + 3 * w4_2), A[3] = 0;
    B[0] = adt * w5_3, B[1] = so + se + w1, B[2] = so - se + w0, B[3] = adt * w6_3;
    C[0] = 0, C[1] = b * (dt2_4 + 3 * w7_2), C[2] = c * (9 - dt2_4 - 12 * w1_2), C[3] = b * (dt2_4 + 3 * w3_2);
  } else if (order == 4) {
    const double a = 1 / 48.0, b = 1 / 24.0, c = 1 / 12.0, d = 1 / 6.0, adt = a * rdt, bdt = b * rdt, cdt = c * rdt;
    double w0 = fabs(delta), w1 = 1 - w0, w2 = 1 - delta, w3 = 1 + delta, w4 = dt - delta, w5 = dt + delta;
    double w0_2 = w0 * w0, w0_3 = w0_2 * w0, w0_4 = w0_3 * w0, w1_2 = w1 * w1, w1_3 = w1_2 * w1;
    double w2_3 = w2 * w2 * w2, w3_3 = w3 * w3 * w3, w4_4 = w4 * w4 * w4 * w4, w5_4 = w5 * w5 * w5 * w5;
    double dt2 = dt * dt, dt3 = dt2 * dt, dt4 = dt3 * dt;
    double ss1 = -dt4 - 6 * w0_2 * dt2 - w0_4;
    double ss2 = 3 * dt4 - 8 * dt3 + 18 * w0_2 * dt2 + (16 - 24 * w0_2) * dt + 3 * w0_4;
    A[0] = d * w0 * (w0_2 + dt2);
    A[1] = d * (4 - 6 * w1_2 + 3 * w1_3 + (1 - 3 * w0) * dt2);
    A[2] = d * (4 - 6 * w0_2 + 3 * w0_3 - (2 - 3 * w0) * dt2);
    A[3] = d * w1 * (w1_2 + dt2);
    A[4] = 0;
    B[0] = adt * w4_4;
    B[1] = cdt * (ss1 + 2 * dt3 * w3 + 2 * dt * (-6 * delta + w3_3));
    B[2] = bdt * ss2;
    B[3] = cdt * (ss1 + 2 * dt3 * w2 + 2 * dt * (+6 * delta + w2_3));
    B[4] = adt * w5_4;
    for (int j = 0; j < 5; j++) C[j] = A[4 - j];
  }
  for (int j = 0; j <= order; j++) s[j] = A[j] * t1 + B[j] * t2 * t3 + C[j] * t4;
}

/* primitives.hpp:158-161 */
double nixo_lorentz_factor(double ux, double uy, double uz, double rc)
{
  return sqrt(1 + (ux * ux + uy * uy + uz * uz) * rc * rc);
}

/* pusher used by the composed step: 0 = Boris, 1 = Vay (2008), 2 = Higuera-Cary (2017); the
 * reference offers the three as interchangeable primitives (primitives.hpp:165-253) */
static int g_pusher = 0;
void nixo_set_pusher(int pusher) { g_pusher = pusher; }
int  nixo_get_pusher(void) { return g_pusher; }

/* primitives.hpp:165-189 */
void nixo_push_boris(double* u, const double* eb, double cc)
{
  double ux = u[0], uy = u[1], uz = u[2];
  double ex = eb[0], ey = eb[1], ez = eb[2], bx = eb[3], by = eb[4], bz = eb[5];
  double gm, bb, vx, vy, vz;

  ux += ex;
  uy += ey;
  uz += ez;

  gm = 1 / sqrt(cc * cc + ux * ux + uy * uy + uz * uz);

  bx *= gm;
  by *= gm;
  bz *= gm;
  bb = 2.0 / (1.0 + bx * bx + by * by + bz * bz);

  vx = ux + (uy * bz - uz * by);
  vy = uy + (uz * bx - ux * bz);
  vz = uz + (ux * by - uy * bx);

  ux += (vy * bz - vz * by) * bb + ex;
  uy += (vz * bx - vx * bz) * bb + ey;
  uz += (vx * by - vy * bx) * bb + ez;

  u[0] = ux;
  u[1] = uy;
  u[2] = uz;
}

/* primitives.hpp:193-224 */
void nixo_push_vay(double* u, const double* eb, double cc)
{
  double ux = u[0], uy = u[1], uz = u[2];
  double ex = eb[0], ey = eb[1], ez = eb[2], bx = eb[3], by = eb[4], bz = eb[5];
  double gm, bb, bu, xx, yy, vx, vy, vz;

  gm = 1 / sqrt(cc * cc + ux * ux + uy * uy + uz * uz);
  vx = ux + 2 * ex + gm * (uy * bz - uz * by);
  vy = uy + 2 * ey + gm * (uz * bx - ux * bz);
  vz = uz + 2 * ez + gm * (ux * by - uy * bx);

  gm = (cc * cc + vx * vx + vy * vy + vz * vz);
  bb = bx * bx + by * by + bz * bz;
  bu = bx * vx + by * vy + bz * vz;
  xx = gm - bb;
  yy = bb + bu * bu;
  gm = 1 / sqrt(0.5 * (xx + sqrt(xx * xx + 4 * yy)));

  bx *= gm;
  by *= gm;
  bz *= gm;
  bu = bx * vx + by * vy + bz * vz;
  bb = 1.0 / (1.0 + bx * bx + by * by + bz * bz);

  u[0] = (vx + bu * bx + (vy * bz - vz * by)) * bb;
  u[1] = (vy + bu * by + (vz * bx - vx * bz)) * bb;
  u[2] = (vz + bu * bz + (vx * by - vy * bx)) * bb;
}

/* primitives.hpp:227-253 */
void nixo_push_higuera_cary(double* u, const double* eb, double cc)
{
  double ux = u[0], uy = u[1], uz = u[2];
  double ex = eb[0], ey = eb[1], ez = eb[2], bx = eb[3], by = eb[4], bz = eb[5];
  double gm, bb, bu, xx, yy, vx, vy, vz;

  ux += ex;
  uy += ey;
  uz += ez;

  gm = cc * cc + ux * ux + uy * uy + uz * uz;
  bb = bx * bx + by * by + bz * bz;
  bu = bx * ux + by * uy + bz * uz;
  xx = gm - bb;
  yy = bb + bu * bu;
  gm = 1 / sqrt(0.5 * (xx + sqrt(xx * xx + 4 * yy)));

  bx *= gm;
  by *= gm;
  bz *= gm;
  bb = 2.0 / (1.0 + bx * bx + by * by + bz * bz);

  vx = ux + (uy * bz - uz * by);
  vy = uy + (uz * bx - ux * bz);
  vz = uz + (ux * by - uy * bx);

  ux += (vy * bz - vz * by) * bb + ex;
  uy += (vz * bx - vx * bz) * bb + ey;
  uz += (vx * by - vy * bx) * bb + ez;
  u[0] = ux;
  u[1] = uy;
  u[2] = uz;
}

/* interp.hpp:149-160 (scalar branch) */
static void interp_shift_weights(int order, int shift, double* ww)
{
  if (shift > 0) {
    for (int ii = order + 1; ii > 0; ii--) {
      ww[ii] = ww[ii - 1];
    }
    ww[0] = 0;
  }
}

void nixo_interp_shift_weights(int order, int shift, double* ww)
{
  interp_shift_weights(order, shift, ww);
}

/* esirkepov.hpp:241-258 (scalar branch): ss is [3][order+3]; a particle that moved one cell down (shift < 0)
 * has its weights moved one slot to the left, one cell up (shift > 0) one slot to the right */
void nixo_esirkepov_shift_weights(int order, const int* shift, double* ss)
{
  const int n = order + 3;
  for (int dir = 0; dir < 3; dir++) {
    double* w = ss + dir * n;
    if (shift[dir] < 0) {
      for (int ii = 0; ii < order + 2; ii++) w[ii] = w[ii + 1];
    } else if (shift[dir] > 0) {
      for (int ii = order + 2; ii > 0; ii--) w[ii] = w[ii - 1];
    }
  }
}

/* interp.hpp:95-113 (interp3d_impl_sorted, scalar) via :217-230 */
double nixo_interp3d(int order, const double* eb, int my, int mx, int iz0, int iy0, int ix0, int ik,
                     const double* wz, const double* wy, const double* wx, double dt)
{
  double result_z = 0;
  for (int jz = 0, iz = iz0; jz < order + 2; jz++, iz++) {
    double result_y = 0;
    for (int jy = 0, iy = iy0; jy < order + 2; jy++, iy++) {
      double result_x = 0;
      for (int jx = 0, ix = ix0; jx < order + 2; jx++, ix++) {
        result_x += eb[(((size_t)iz * my + iy) * mx + ix) * 6 + ik] * wx[jx];
      }
      result_y += result_x * wy[jy];
    }
    result_z += result_y * wz[jz];
  }
  return result_z * dt;
}

/* esirkepov.hpp:154-237 + 325-340.  ss is [2][3][n], cur is [n][n][n][4], n = order+3 */
void nixo_deposit3d(int order, double dxdt, double dydt, double dzdt, double qs, double* ss,
                    double* cur)
{
  const int    n = order + 3;
  const double A = 1.0 / 2;
  const double B = 1.0 / 3;
#define SS(t, d, j) ss[((t) * 3 + (d)) * n + (j)]
#define CUR(z, y, x, k) cur[((((z) * n + (y)) * n + (x)) * 4) + (k)]

  /* ro3d :155-164 */
  for (int jz = 0; jz < n; jz++)
    for (int jy = 0; jy < n; jy++)
      for (int jx = 0; jx < n; jx++)
        CUR(jz, jy, jx, 0) += qs * SS(1, 0, jx) * SS(1, 1, jy) * SS(1, 2, jz);

  /* ds3d :167-174 */
  for (int dir = 0; dir < 3; dir++)
    for (int l = 0; l < n; l++)
      SS(1, dir, l) -= SS(0, dir, l);

  /* jx3d :177-195 */
  {
    double qdxdt = qs * dxdt;
    for (int jz = 0; jz < n; jz++)
      for (int jy = 0; jy < n; jy++) {
        double ww = 0;
        double wx = -((1 * SS(0, 1, jy) + A * SS(1, 1, jy)) * SS(0, 2, jz) +
                      (A * SS(0, 1, jy) + B * SS(1, 1, jy)) * SS(1, 2, jz)) *
                    qdxdt;
        for (int jx = 0; jx < n - 1; jx++) {
          ww += SS(1, 0, jx) * wx;
          CUR(jz, jy, jx + 1, 1) += ww;
        }
      }
  }
  /* jy3d :198-216 */
  {
    double qdydt = qs * dydt;
    for (int jz = 0; jz < n; jz++)
      for (int jx = 0; jx < n; jx++) {
        double ww = 0;
        double wy = -((1 * SS(0, 2, jz) + A * SS(1, 2, jz)) * SS(0, 0, jx) +
                      (A * SS(0, 2, jz) + B * SS(1, 2, jz)) * SS(1, 0, jx)) *
                    qdydt;
        for (int jy = 0; jy < n - 1; jy++) {
          ww += SS(1, 1, jy) * wy;
          CUR(jz, jy + 1, jx, 2) += ww;
        }
      }
  }
  /* jz3d :219-237 */
  {
    double qdzdt = qs * dzdt;
    for (int jy = 0; jy < n; jy++)
      for (int jx = 0; jx < n; jx++) {
        double ww = 0;
        double wz = -((1 * SS(0, 0, jx) + A * SS(1, 0, jx)) * SS(0, 1, jy) +
                      (A * SS(0, 0, jx) + B * SS(1, 0, jx)) * SS(1, 1, jy)) *
                    qdzdt;
        for (int jz = 0; jz < n - 1; jz++) {
          ww += SS(1, 2, jz) * wz;
          CUR(jz + 1, jy, jx, 3) += ww;
        }
      }
  }
#undef SS
#undef CUR
}

/* append_current3d<order>, scalar branch (primitives.hpp:786-797): uj is [mz][my][mx][4], cur [n][n][n][4],
 * n = order + 3 */
void nixo_append_current3d(int order, double* uj, int my, int mx, int iz0, int iy0, int ix0, const double* cur)
{
  const int n = order + 3;
  for (int jz = 0, iz = iz0; jz < n; jz++, iz++)
    for (int jy = 0, iy = iy0; jy < n; jy++, iy++)
      for (int jx = 0, ix = ix0; jx < n; jx++, ix++)
        for (int k = 0; k < 4; k++)
          uj[(((size_t)iz * my + iy) * mx + ix) * 4 + k] += cur[(((jz * n) + jy) * n + jx) * 4 + k];
}

/* append_moment3d<order>, scalar branch (primitives.hpp:904-915): um is [mz][my][mx][ns][14], mom [n][n][n][14],
 * n = order + 1 */
void nixo_append_moment3d(int order, double* um, int my, int mx, int ns, int iz0, int iy0, int ix0, int is,
                          const double* mom)
{
  const int n = order + 1;
  for (int jz = 0, iz = iz0; jz < n; jz++, iz++)
    for (int jy = 0, iy = iy0; jy < n; jy++, iy++)
      for (int jx = 0, ix = ix0; jx < n; jx++, ix++)
        for (int k = 0; k < 14; k++)
          um[((((size_t)iz * my + iy) * mx + ix) * ns + is) * 14 + k] += mom[(((jz * n) + jy) * n + jx) * 14 + k];
}

/* ------------------------------------------------------------------------------------------ */
/* chunk geometry                                                                             */
/* ------------------------------------------------------------------------------------------ */

/* chunk.cpp:118-208 (all three dimensions present) */
static void set_index_bounds(nixo_chunk* c)
{
  int nb = c->g.nb;
  for (int a = 0; a < 3; a++) {
    int Lb       = nb;
    int Ub       = nb + c->g.dims[a] - 1;
    c->Lb[a]     = Lb;
    c->Ub[a]     = Ub;
    c->sendlb[a][0] = Lb;
    c->sendlb[a][1] = Lb;
    c->sendlb[a][2] = Ub - nb + 1;
    c->sendub[a][0] = Lb + nb - 1;
    c->sendub[a][1] = Ub;
    c->sendub[a][2] = Ub;
    c->recvlb[a][0] = Lb - nb;
    c->recvlb[a][1] = Lb;
    c->recvlb[a][2] = Ub + 1;
    c->recvub[a][0] = Lb - 1;
    c->recvub[a][1] = Ub;
    c->recvub[a][2] = Ub + nb;
  }
}

/* chunk.cpp:210-237 */
static void set_coordinate(nixo_chunk* c)
{
  for (int a = 0; a < 3; a++) {
    double del   = c->g.del[a];
    c->lim[a][0]  = c->g.offset[a] * del;
    c->lim[a][1]  = c->g.offset[a] * del + c->g.dims[a] * del;
    c->glim[a][0] = 0.0;
    c->glim[a][1] = c->g.gdims[a] * del;
  }
}

/* chunk.cpp:257-286 */
static void set_mpi_buffer(nixo_chunk* c, mpibuf_t* mb, int headbyte, int elembyte)
{
  int size = 0;
  for (int iz = 0; iz <= 2; iz++)
    for (int iy = 0; iy <= 2; iy++)
      for (int ix = 0; ix <= 2; ix++) {
        int s = 9 * iz + 3 * iy + ix;
        if (iz == 1 && iy == 1 && ix == 1) {
          mb->bufsize[s] = 0;
          mb->bufaddr[s] = size;
        } else {
          int nz         = c->recvub[0][iz] - c->recvlb[0][iz] + 1;
          int ny         = c->recvub[1][iy] - c->recvlb[1][iy] + 1;
          int nx         = c->recvub[2][ix] - c->recvlb[2][ix] + 1;
          mb->bufsize[s] = headbyte + elembyte * nz * ny * nx;
          mb->bufaddr[s] = size;
          size += mb->bufsize[s];
        }
      }
  mb->sendbuf  = (uint8_t*)calloc((size_t)size + 1, 1);
  mb->recvbuf  = (uint8_t*)calloc((size_t)size + 1, 1);
  mb->sendsize = size;
  mb->recvsize = size;
}

/* buffer.hpp:41-53 */
static void buffer_resize(uint8_t** buf, int* size, int s)
{
  int      copysize = (*size < s) ? *size : s;
  uint8_t* p        = (uint8_t*)calloc((size_t)s + 1, 1);
  memcpy(p, *buf, (size_t)copysize);
  free(*buf);
  *buf  = p;
  *size = s;
}

/* particle.hpp:146-153 */
static int round_up_alloc(int np_required)
{
  return ((np_required + ALLOC_UNIT) / ALLOC_UNIT) * ALLOC_UNIT;
}

/* particle.hpp:80-107 + xtensor_particle.hpp:49-65 */
static void particle_init(nixo_chunk* c, particle_t* p, int np_required, double q, double m)
{
  p->Np       = 0;
  p->Ng       = c->M[0] * c->M[1] * c->M[2];
  p->q        = q;
  p->m        = m;
  p->Np_total = round_up_alloc(np_required);
  p->xu       = (double*)calloc((size_t)p->Np_total * NC, sizeof(double));
  p->xv       = (double*)calloc((size_t)p->Np_total * NC, sizeof(double));
  p->gindex   = (int32_t*)calloc((size_t)p->Np_total, sizeof(int32_t));
  p->pindex   = (int32_t*)calloc((size_t)p->Ng + 1, sizeof(int32_t));
  p->pcount   = (int32_t*)calloc(((size_t)p->Ng + 1) * LANES, sizeof(int32_t));
}

nixo_chunk* nixo_chunk_create(const nixo_geom* g, int ns, const int* np_required, const double* q,
                              const double* m)
{
  nixo_chunk* c = (nixo_chunk*)calloc(1, sizeof(nixo_chunk));
  c->g          = *g;
  c->ns         = ns;
  set_index_bounds(c);
  set_coordinate(c);
  for (int a = 0; a < 3; a++)
    c->M[a] = g->dims[a] + 2 * g->nb;
  for (int s = 0; s < 27; s++)
    c->nbvalid[s] = 1;
  size_t ncell = (size_t)c->M[0] * c->M[1] * c->M[2];
  c->uf        = (double*)calloc(ncell * 6, sizeof(double));
  c->uj        = (double*)calloc(ncell * 4, sizeof(double));
  c->um        = (double*)calloc(ncell * (size_t)ns * 14, sizeof(double));
  c->up        = (particle_t*)calloc((size_t)ns, sizeof(particle_t));
  for (int is = 0; is < ns; is++)
    particle_init(c, &c->up[is], np_required[is], q[is], m[is]);
  set_mpi_buffer(c, &c->mpibuf[NIXO_MODE_FIELD], 0, 8 * 6);
  set_mpi_buffer(c, &c->mpibuf[NIXO_MODE_CURRENT], 0, 8 * 4);
  set_mpi_buffer(c, &c->mpibuf[NIXO_MODE_PARTICLE], HEAD_BYTE, ELEM_BYTE);
  set_mpi_buffer(c, &c->mpibuf[NIXO_MODE_MOMENT], 0, 8 * ns * 14);
  c->num_unpacked = (int*)calloc((size_t)ns, sizeof(int));
  return c;
}

void nixo_chunk_destroy(nixo_chunk* c)
{
  if (!c)
    return;
  for (int is = 0; is < c->ns; is++) {
    free(c->up[is].xu);
    free(c->up[is].xv);
    free(c->up[is].gindex);
    free(c->up[is].pindex);
    free(c->up[is].pcount);
  }
  for (int mode = 0; mode < 4; mode++) {
    free(c->mpibuf[mode].sendbuf);
    free(c->mpibuf[mode].recvbuf);
  }
  free(c->up);
  free(c->uf);
  free(c->uj);
  free(c->um);
  free(c->num_unpacked);
  free(c);
}

double* nixo_chunk_uf(nixo_chunk* c)
{
  return c->uf;
}
double* nixo_chunk_uj(nixo_chunk* c)
{
  return c->uj;
}
void nixo_chunk_set_nb_valid(nixo_chunk* c, int iz, int iy, int ix, int valid)
{
  c->nbvalid[9 * iz + 3 * iy + ix] = valid;
}

/* ------------------------------------------------------------------------------------------ */
/* particle container                                                                         */
/* ------------------------------------------------------------------------------------------ */
int nixo_particle_ng(nixo_chunk* c, int is)
{
  return c->up[is].Ng;
}
int nixo_particle_np(nixo_chunk* c, int is)
{
  return c->up[is].Np;
}
void nixo_particle_set_np(nixo_chunk* c, int is, int np)
{
  c->up[is].Np = np;
}
int nixo_particle_np_total(nixo_chunk* c, int is)
{
  return c->up[is].Np_total;
}
double* nixo_particle_xu(nixo_chunk* c, int is)
{
  return c->up[is].xu;
}
double* nixo_particle_xv(nixo_chunk* c, int is)
{
  return c->up[is].xv;
}
int32_t* nixo_particle_gindex(nixo_chunk* c, int is)
{
  return c->up[is].gindex;
}
int32_t* nixo_particle_pindex(nixo_chunk* c, int is)
{
  return c->up[is].pindex;
}
int32_t* nixo_particle_pcount(nixo_chunk* c, int is)
{
  return c->up[is].pcount;
}

/* xtensor_particle.hpp:70-115 */
void nixo_particle_resize(nixo_chunk* c, int is, int np_required)
{
  particle_t* p      = &c->up[is];
  int         np_new = round_up_alloc(np_required);
  if (np_new == p->Np_total || np_new <= p->Np)
    return;
  size_t ncopy = (size_t)(p->Np_total < np_new ? p->Np_total : np_new);
  double*  xu  = (double*)calloc((size_t)np_new * NC, sizeof(double));
  double*  xv  = (double*)calloc((size_t)np_new * NC, sizeof(double));
  int32_t* gi  = (int32_t*)calloc((size_t)np_new, sizeof(int32_t));
  memcpy(xu, p->xu, ncopy * NC * sizeof(double));
  memcpy(xv, p->xv, ncopy * NC * sizeof(double));
  memcpy(gi, p->gindex, ncopy * sizeof(int32_t));
  free(p->xu);
  free(p->xv);
  free(p->gindex);
  p->xu       = xu;
  p->xv       = xv;
  p->gindex   = gi;
  p->Np_total = np_new;
}

/* xtensor_particle.hpp:231-238 */
static int flatindex(const nixo_chunk* c, int iz, int iy, int ix)
{
  const int stride_x = 1;
  const int stride_y = stride_x * (c->Ub[2] - c->Lb[2] + 2);
  const int stride_z = stride_y * (c->Ub[1] - c->Lb[1] + 2);
  return iz * stride_z + iy * stride_y + ix * stride_x;
}

/* xtensor_particle.hpp:324-357 (+ increment :246-252) */
void nixo_particle_count(nixo_chunk* c, int is, int lbp, int ubp, int reset, int order)
{
  particle_t*  p             = &c->up[is];
  const int    is_odd        = (order % 2 == 1) ? 1 : 0;
  const int    out_of_bounds = p->Ng;
  const double delx = c->g.del[2], dely = c->g.del[1], delz = c->g.del[0];
  const double xmin = c->lim[2][0], xmax = c->lim[2][1];
  const double ymin = c->lim[1][0], ymax = c->lim[1][1];
  const double zmin = c->lim[0][0], zmax = c->lim[0][1];
  const double xoffset = xmin - 0.5 * delx * is_odd;
  const double yoffset = ymin - 0.5 * dely * is_odd;
  const double zoffset = zmin - 0.5 * delz * is_odd;
  const double rdx     = 1 / delx;
  const double rdy     = 1 / dely;
  const double rdz     = 1 / delz;

  if (reset) {
    memset(p->pcount, 0, sizeof(int32_t) * LANES * ((size_t)p->Ng + 1));
  }

  for (int ip = lbp; ip <= ubp; ip++) {
    const double* xu = &p->xu[(size_t)ip * NC];
    int           ix = nixo_digitize(xu[0], xoffset, rdx);
    int           iy = nixo_digitize(xu[1], yoffset, rdy);
    int           iz = nixo_digitize(xu[2], zoffset, rdz);
    int           ii = flatindex(c, iz, iy, ix);

    ii = (xu[0] < xmin || xu[0] >= xmax) ? out_of_bounds : ii;
    ii = (xu[1] < ymin || xu[1] >= ymax) ? out_of_bounds : ii;
    ii = (xu[2] < zmin || xu[2] >= zmax) ? out_of_bounds : ii;

    int jj        = ip % LANES;
    p->gindex[ip] = ii;
    p->pcount[(size_t)ii * LANES + jj]++;
  }
}

/* xtensor_particle.hpp:260-321 */
void nixo_particle_sort(nixo_chunk* c, int is)
{
  particle_t* p  = &c->up[is];
  const int   Ng = p->Ng;
#define PC(ii, jj) p->pcount[(size_t)(ii) * LANES + (jj)]
  for (int ii = 0; ii < Ng + 1; ii++)
    for (int jj = 0; jj < LANES - 1; jj++)
      PC(ii, jj + 1) += PC(ii, jj);
  for (int ii = 0; ii < Ng; ii++)
    for (int jj = 0; jj < LANES; jj++)
      PC(ii + 1, jj) += PC(ii, LANES - 1);

  p->pindex[0] = 0;
  for (int ii = 0; ii < Ng; ii++)
    p->pindex[ii + 1] = PC(ii, LANES - 1);

  for (int ii = 0; ii < Ng + 1; ii++)
    for (int jj = LANES - 1; jj > 0; jj--)
      PC(ii, jj) = PC(ii, jj - 1);
  for (int ii = 0; ii < Ng + 1; ii++)
    PC(ii, 0) = p->pindex[ii];

  for (int ip = 0; ip < p->Np; ip++) {
    int ii = p->gindex[ip];
    int jj = ip % LANES;
    int jp = PC(ii, jj);
    memcpy(&p->xv[(size_t)NC * jp], &p->xu[(size_t)NC * ip], NC * sizeof(double));
    PC(ii, jj)++;
  }
#undef PC
  /* swap (:120-123) */
  double* t = p->xu;
  p->xu     = p->xv;
  p->xv     = t;

  p->Np = p->pindex[Ng];
}

/* xtensor_particle.hpp:359-376 */
void nixo_particle_set_boundary_periodic(nixo_chunk* c, int is, int lbp, int ubp)
{
  particle_t*  p  = &c->up[is];
  const double X1 = c->glim[2][0], X2 = c->glim[2][1];
  const double Y1 = c->glim[1][0], Y2 = c->glim[1][1];
  const double Z1 = c->glim[0][0], Z2 = c->glim[0][1];
  const double X  = 1 * (X2 - X1);
  const double Y  = 1 * (Y2 - Y1);
  const double Z  = 1 * (Z2 - Z1);
  for (int ip = lbp; ip <= ubp; ip++) {
    double* xu = &p->xu[(size_t)ip * NC];
    xu[0] += (xu[0] < X1) * X - (xu[0] >= X2) * X;
    xu[1] += (xu[1] < Y1) * Y - (xu[1] >= Y2) * Y;
    xu[2] += (xu[2] < Z1) * Z - (xu[2] >= Z2) * Z;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* composed step (a)+(b): the same composition as oracle/ref/ref_driver.cpp push_deposit_scalar */
/* ------------------------------------------------------------------------------------------ */
static void push_deposit_one(nixo_chunk* c, particle_t* p, int ip, double delt, double cc)
{
  const int order  = c->g.order;
  const int is_odd = order % 2;
  const int half   = order / 2;
  const int n      = order + 3;
  const int Lbx = c->Lb[2], Lby = c->Lb[1], Lbz = c->Lb[0];
  const double delx = c->g.del[2], dely = c->g.del[1], delz = c->g.del[0];
  const double xmin = c->lim[2][0], ymin = c->lim[1][0], zmin = c->lim[0][0];
  const double rdx = 1 / delx, rdy = 1 / dely, rdz = 1 / delz;
  const double rc   = 1 / cc;
  const double dt1  = 0.5 * p->q / p->m * delt;
  const double dxdt = delx / delt, dydt = dely / delt, dzdt = delz / delt;
  const double xoff  = xmin - 0.5 * delx * is_odd;
  const double yoff  = ymin - 0.5 * dely * is_odd;
  const double zoff  = zmin - 0.5 * delz * is_odd;
  const double xhoff = xmin - 0.5 * delx * (1 - is_odd);
  const double yhoff = ymin - 0.5 * dely * (1 - is_odd);
  const double zhoff = zmin - 0.5 * delz * (1 - is_odd);
  const double ximin = xmin + 0.5 * delx;
  const double yimin = ymin + 0.5 * dely;
  const double zimin = zmin + 0.5 * delz;
  const int    my = c->M[1], mx = c->M[2];

  double* xu = &p->xu[(size_t)ip * NC];
  double* xv = &p->xv[(size_t)ip * NC];
  double  x = xu[0], y = xu[1], z = xu[2];
  double  u[3] = {xu[3], xu[4], xu[5]};

  int ix = nixo_digitize(x, xoff, rdx) - is_odd;
  int iy = nixo_digitize(y, yoff, rdy) - is_odd;
  int iz = nixo_digitize(z, zoff, rdz) - is_odd;
  int hx = nixo_digitize(x, xhoff, rdx);
  int hy = nixo_digitize(y, yhoff, rdy);
  int hz = nixo_digitize(z, zhoff, rdz);

  double wix[MAXO + 2] = {0}, wiy[MAXO + 2] = {0}, wiz[MAXO + 2] = {0};
  double whx[MAXO + 2] = {0}, why[MAXO + 2] = {0}, whz[MAXO + 2] = {0};
  nixo_shape_mc(order, x, ximin + ix * delx, rdx, wix);
  nixo_shape_mc(order, y, yimin + iy * dely, rdy, wiy);
  nixo_shape_mc(order, z, zimin + iz * delz, rdz, wiz);
  nixo_shape_mc(order, x, xmin + hx * delx, rdx, whx);
  nixo_shape_mc(order, y, ymin + hy * dely, rdy, why);
  nixo_shape_mc(order, z, zmin + hz * delz, rdz, whz);

  int ix0 = ix - half + Lbx, iy0 = iy - half + Lby, iz0 = iz - half + Lbz;
  interp_shift_weights(order, hx - ix, whx);
  interp_shift_weights(order, hy - iy, why);
  interp_shift_weights(order, hz - iz, whz);

  double eb[6];
  eb[0] = nixo_interp3d(order, c->uf, my, mx, iz0, iy0, ix0, 0, wiz, wiy, whx, dt1);
  eb[1] = nixo_interp3d(order, c->uf, my, mx, iz0, iy0, ix0, 1, wiz, why, wix, dt1);
  eb[2] = nixo_interp3d(order, c->uf, my, mx, iz0, iy0, ix0, 2, whz, wiy, wix, dt1);
  eb[3] = nixo_interp3d(order, c->uf, my, mx, iz0, iy0, ix0, 3, whz, why, wix, dt1);
  eb[4] = nixo_interp3d(order, c->uf, my, mx, iz0, iy0, ix0, 4, whz, wiy, whx, dt1);
  eb[5] = nixo_interp3d(order, c->uf, my, mx, iz0, iy0, ix0, 5, wiz, why, whx, dt1);

  if (g_pusher == 1) nixo_push_vay(u, eb, cc);
  else if (g_pusher == 2) nixo_push_higuera_cary(u, eb, cc);
  else nixo_push_boris(u, eb, cc);

  double gam = nixo_lorentz_factor(u[0], u[1], u[2], rc);
  double dtg = delt / gam;

  xv[0] = x;
  xv[1] = y;
  xv[2] = z;
  xu[0] = x + u[0] * dtg;
  xu[1] = y + u[1] * dtg;
  xu[2] = z + u[2] * dtg;
  xu[3] = u[0];
  xu[4] = u[1];
  xu[5] = u[2];

  double ss[2][3][MAXO + 3];
  double ssn[2 * 3 * (MAXO + 3)];
  memset(ss, 0, sizeof(ss));
  nixo_shape_mc(order, xv[0], ximin + ix * delx, rdx, &ss[0][0][1]);
  nixo_shape_mc(order, xv[1], yimin + iy * dely, rdy, &ss[0][1][1]);
  nixo_shape_mc(order, xv[2], zimin + iz * delz, rdz, &ss[0][2][1]);

  int ix1 = nixo_digitize(xu[0], xoff, rdx) - is_odd;
  int iy1 = nixo_digitize(xu[1], yoff, rdy) - is_odd;
  int iz1 = nixo_digitize(xu[2], zoff, rdz) - is_odd;
  if (abs(ix1 - ix) > 1 || abs(iy1 - iy) > 1 || abs(iz1 - iz) > 1) {
    return;
  }
  nixo_shape_mc(order, xu[0], ximin + ix1 * delx, rdx, &ss[1][0][1 + ix1 - ix]);
  nixo_shape_mc(order, xu[1], yimin + iy1 * dely, rdy, &ss[1][1][1 + iy1 - iy]);
  nixo_shape_mc(order, xu[2], zimin + iz1 * delz, rdz, &ss[1][2][1 + iz1 - iz]);

  /* repack ss to the dense [2][3][n] layout of the order at hand */
  for (int t = 0; t < 2; t++)
    for (int d = 0; d < 3; d++)
      for (int j = 0; j < n; j++)
        ssn[(t * 3 + d) * n + j] = ss[t][d][j];

  double cur[(MAXO + 3) * (MAXO + 3) * (MAXO + 3) * 4];
  memset(cur, 0, sizeof(double) * (size_t)n * n * n * 4);
  nixo_deposit3d(order, dxdt, dydt, dzdt, p->q, ssn, cur);

  /* append_current3d scalar branch, primitives.hpp:786-797 */
  int jx0 = ix - half - 1 + Lbx, jy0 = iy - half - 1 + Lby, jz0 = iz - half - 1 + Lbz;
  for (int jz = 0, kz = jz0; jz < n; jz++, kz++)
    for (int jy = 0, ky = jy0; jy < n; jy++, ky++)
      for (int jx = 0, kx = jx0; jx < n; jx++, kx++) {
        double*       dst = &c->uj[(((size_t)kz * my + ky) * mx + kx) * 4];
        const double* src = &cur[(((jz * n) + jy) * n + jx) * 4];
        dst[0] += src[0];
        dst[1] += src[1];
        dst[2] += src[2];
        dst[3] += src[3];
      }
}

void nixo_chunk_push_deposit(nixo_chunk* c, double delt, double cc, int simd)
{
  (void)simd; /* the restatement has only the scalar instantiation */
  for (int is = 0; is < c->ns; is++) {
    particle_t* p = &c->up[is];
    for (int ip = 0; ip < p->Np; ip++)
      push_deposit_one(c, p, ip, delt, cc);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* halo engine                                                                                */
/* ------------------------------------------------------------------------------------------ */
static void get_bounds(const nixo_chunk* c, int iz, int iy, int ix, int use_recv, int lo[3],
                       int hi[3])
{
  const int d[3] = {iz, iy, ix};
  for (int a = 0; a < 3; a++) {
    lo[a] = use_recv ? c->recvlb[a][d[a]] : c->sendlb[a][d[a]];
    hi[a] = use_recv ? c->recvub[a][d[a]] : c->sendub[a][d[a]];
  }
}

/* slab <-> contiguous buffer in (z,y,x,c) row-major order: std::copy over xt::strided_view
 * (xtensor_halo3d.hpp:35-41,61-67,93-99,119-125).  op: 0 = slab->buf, 1 = buf->slab, 2 = slab+=buf */
static void slab_copy(const nixo_chunk* c, double* data, int ncomp, const int lo[3], const int hi[3],
                      double* buf, int op)
{
  size_t k = 0;
  for (int iz = lo[0]; iz <= hi[0]; iz++)
    for (int iy = lo[1]; iy <= hi[1]; iy++)
      for (int ix = lo[2]; ix <= hi[2]; ix++) {
        double* cell = &data[(((size_t)iz * c->M[1] + iy) * c->M[2] + ix) * ncomp];
        for (int ic = 0; ic < ncomp; ic++, k++) {
          if (op == 0)
            buf[k] = cell[ic];
          else if (op == 1)
            cell[ic] = buf[k];
          else
            cell[ic] = buf[k] + cell[ic]; /* std::plus(ptr[i], view[i]) : :125 */
        }
      }
}

/* XtensorHaloParticle3D::pre_pack, xtensor_halo3d.hpp:273-406 */
static void particle_pre_pack(nixo_chunk* c)
{
  mpibuf_t* mb = &c->mpibuf[NIXO_MODE_PARTICLE];
  const int Ns = c->ns;
  const double xmin = c->lim[2][0], xmax = c->lim[2][1];
  const double ymin = c->lim[1][0], ymax = c->lim[1][1];
  const double zmin = c->lim[0][0], zmax = c->lim[0][1];
  int* send_count = (int*)calloc((size_t)(Ns + 1) * 27, sizeof(int));

  for (int is = 0; is < Ns; is++) {
    particle_t* p = &c->up[is];
    for (int ip = 0; ip < p->Np; ip++) {
      const double* xu = &p->xu[(size_t)ip * NC];
      int iz = (xu[2] >= zmax) - (xu[2] < zmin) + 1;
      int iy = (xu[1] >= ymax) - (xu[1] < ymin) + 1;
      int ix = (xu[0] >= xmax) - (xu[0] < xmin) + 1;
      if (ix == 1 && iy == 1 && iz == 1)
        continue;
      send_count[is * 27 + 9 * iz + 3 * iy + ix]++;
      send_count[Ns * 27 + 9 * iz + 3 * iy + ix]++;
    }
  }

  int bufsize = 0;
  for (int s = 0; s < 27; s++) {
    mb->bufsize[s] = 0;
    mb->bufaddr[s] = 0;
  }
  for (int s = 0; s < 27; s++) {
    if (s == 13)
      continue;
    mb->bufsize[s] = ELEM_BYTE * send_count[Ns * 27 + s] + HEAD_BYTE * Ns;
    mb->bufaddr[s] = bufsize;
    bufsize += mb->bufsize[s];
  }
  buffer_resize(&mb->sendbuf, &mb->sendsize, bufsize);

  /* headers */
  for (int s = 0; s < 27; s++) {
    if (s == 13)
      continue;
    int addr = mb->bufaddr[s];
    for (int is = 0; is < Ns; is++) {
      memcpy(mb->sendbuf + addr, &send_count[is * 27 + s], HEAD_BYTE);
      addr += HEAD_BYTE + ELEM_BYTE * send_count[is * 27 + s];
    }
  }

  /* payload, species-major inside each direction */
  int addr[27];
  memcpy(addr, mb->bufaddr, sizeof(addr));
  for (int is = 0; is < Ns; is++) {
    for (int s = 0; s < 27; s++)
      addr[s] += HEAD_BYTE;
    particle_t* p = &c->up[is];
    for (int ip = 0; ip < p->Np; ip++) {
      const double* xu = &p->xu[(size_t)ip * NC];
      int iz = (xu[2] >= zmax) - (xu[2] < zmin) + 1;
      int iy = (xu[1] >= ymax) - (xu[1] < ymin) + 1;
      int ix = (xu[0] >= xmax) - (xu[0] < xmin) + 1;
      if (ix == 1 && iy == 1 && iz == 1)
        continue;
      int s = 9 * iz + 3 * iy + ix;
      memcpy(mb->sendbuf + addr[s], xu, ELEM_BYTE);
      addr[s] += ELEM_BYTE;
    }
  }
  free(send_count);
}

/* XtensorHaloParticle3D::pre_unpack, xtensor_halo3d.hpp:429-482 */
static void particle_pre_unpack(nixo_chunk* c)
{
  mpibuf_t* mb = &c->mpibuf[NIXO_MODE_PARTICLE];
  const int Ns = c->ns;
  int* recv_count = (int*)calloc((size_t)(Ns + 1) * 27, sizeof(int));
  for (int s = 0; s < 27; s++) {
    if (s == 13 || !c->nbvalid[s])
      continue;
    int rcnt = 0;
    int addr = mb->bufaddr[s];
    for (int is = 0; is < Ns; is++) {
      memcpy(&rcnt, mb->recvbuf + addr, HEAD_BYTE);
      addr += HEAD_BYTE + ELEM_BYTE * rcnt;
      recv_count[is * 27 + s] = rcnt;
      recv_count[Ns * 27 + s] += rcnt;
    }
  }
  for (int is = 0; is < Ns; is++) {
    int np_next = c->up[is].Np;
    for (int s = 0; s < 27; s++)
      np_next += recv_count[is * 27 + s];
    nixo_particle_resize(c, is, np_next);
    c->num_unpacked[is] = 0;
  }
  free(recv_count);
}

/* XtensorHaloParticle3D::unpack, xtensor_halo3d.hpp:485-533 */
static void particle_unpack(nixo_chunk* c, int s)
{
  mpibuf_t* mb      = &c->mpibuf[NIXO_MODE_PARTICLE];
  uint8_t*  recvptr = mb->recvbuf + mb->bufaddr[s];
  int       recvcnt = mb->bufsize[s];
  if (recvcnt < c->ns * HEAD_BYTE)
    return;
  for (int is = 0; is < c->ns; is++) {
    particle_t* p = &c->up[is];
    int         rcnt;
    memcpy(&rcnt, recvptr, HEAD_BYTE);
    recvptr += HEAD_BYTE;
    recvcnt -= HEAD_BYTE;
    memcpy(&p->xu[(size_t)(p->Np + c->num_unpacked[is]) * NC], recvptr, (size_t)ELEM_BYTE * rcnt);
    recvptr += rcnt * ELEM_BYTE;
    recvcnt -= rcnt * ELEM_BYTE;
    c->num_unpacked[is] += rcnt;
  }
}

/* XtensorHaloParticle3D::post_unpack, xtensor_halo3d.hpp:536-556 */
static void particle_post_unpack(nixo_chunk* c)
{
  for (int is = 0; is < c->ns; is++) {
    particle_t* p       = &c->up[is];
    int         np_prev = p->Np;
    int         np_next = p->Np + c->num_unpacked[is];
    nixo_particle_set_boundary_periodic(c, is, np_prev, np_next - 1);
    nixo_particle_count(c, is, np_prev, np_next - 1, 0, c->g.order);
    p->Np = np_next;
  }
  for (int is = 0; is < c->ns; is++)
    nixo_particle_sort(c, is);
}

/* Chunk::pack_bc_exchange (chunk.hpp:435-468) with the Halo class selected by `mode` */
void nixo_chunk_halo_pack(nixo_chunk* c, int mode)
{
  mpibuf_t* mb = &c->mpibuf[mode];
  int       lo[3], hi[3];
  if (mode == NIXO_MODE_PARTICLE) {
    particle_pre_pack(c);
    return;
  }
  for (int iz = 0; iz <= 2; iz++)
    for (int iy = 0; iy <= 2; iy++)
      for (int ix = 0; ix <= 2; ix++) {
        int s = 9 * iz + 3 * iy + ix;
        if (s == 13)
          continue;
        double* buf = (double*)(mb->sendbuf + mb->bufaddr[s]);
        if (mode == NIXO_MODE_FIELD) {
          get_bounds(c, iz, iy, ix, 0, lo, hi); /* send slab (interior), :35-41 */
          slab_copy(c, c->uf, 6, lo, hi, buf, 0);
        } else if (mode == NIXO_MODE_MOMENT) {
          get_bounds(c, iz, iy, ix, 1, lo, hi); /* recv slab (ghost), XtensorHaloMoment3D::pack :144-160 */
          slab_copy(c, c->um, c->ns * 14, lo, hi, buf, 0);
        } else {
          get_bounds(c, iz, iy, ix, 1, lo, hi); /* recv slab (ghost), :93-99 */
          slab_copy(c, c->uj, 4, lo, hi, buf, 0);
        }
      }
}

/* Chunk::unpack_bc_exchange (chunk.hpp:471-504) */
void nixo_chunk_halo_unpack(nixo_chunk* c, int mode)
{
  mpibuf_t* mb = &c->mpibuf[mode];
  int       lo[3], hi[3];
  if (mode == NIXO_MODE_PARTICLE)
    particle_pre_unpack(c);
  for (int iz = 0; iz <= 2; iz++)
    for (int iy = 0; iy <= 2; iy++)
      for (int ix = 0; ix <= 2; ix++) {
        int s = 9 * iz + 3 * iy + ix;
        if (s == 13)
          continue;
        if (!c->nbvalid[s]) /* MPI_PROC_NULL neighbour, :57,115,491 */
          continue;
        double* buf = (double*)(mb->recvbuf + mb->bufaddr[s]);
        if (mode == NIXO_MODE_FIELD) {
          get_bounds(c, iz, iy, ix, 1, lo, hi); /* recv slab (ghost), :61-67 */
          slab_copy(c, c->uf, 6, lo, hi, buf, 1);
        } else if (mode == NIXO_MODE_CURRENT) {
          get_bounds(c, iz, iy, ix, 0, lo, hi); /* send slab (interior) +=, :119-125 */
          slab_copy(c, c->uj, 4, lo, hi, buf, 2);
        } else if (mode == NIXO_MODE_MOMENT) {
          get_bounds(c, iz, iy, ix, 0, lo, hi); /* send slab (interior) +=, XtensorHaloMoment3D::unpack :163-186 */
          slab_copy(c, c->um, c->ns * 14, lo, hi, buf, 2);
        } else {
          particle_unpack(c, s);
        }
      }
  if (mode == NIXO_MODE_PARTICLE)
    particle_post_unpack(c);
}

int nixo_chunk_bufsize(nixo_chunk* c, int mode, int iz, int iy, int ix)
{
  return c->mpibuf[mode].bufsize[9 * iz + 3 * iy + ix];
}
int nixo_chunk_bufaddr(nixo_chunk* c, int mode, int iz, int iy, int ix)
{
  return c->mpibuf[mode].bufaddr[9 * iz + 3 * iy + ix];
}
uint8_t* nixo_chunk_sendbuf(nixo_chunk* c, int mode)
{
  return c->mpibuf[mode].sendbuf;
}
int nixo_chunk_sendbuf_size(nixo_chunk* c, int mode)
{
  return c->mpibuf[mode].sendsize;
}

/* probe_bc_exchange once every message is ready, chunk.cpp:356-368 */
void nixo_chunk_set_recv_sizes(nixo_chunk* c, int mode, const int* bufsize27)
{
  mpibuf_t* mb      = &c->mpibuf[mode];
  int       bufsize = 0;
  for (int s = 0; s < 27; s++) {
    mb->bufsize[s] = bufsize27[s];
    mb->bufaddr[s] = bufsize;
    bufsize += mb->bufsize[s];
  }
  buffer_resize(&mb->recvbuf, &mb->recvsize, bufsize);
}
uint8_t* nixo_chunk_recvbuf(nixo_chunk* c, int mode)
{
  return c->mpibuf[mode].recvbuf;
}
int nixo_chunk_recvbuf_size(nixo_chunk* c, int mode)
{
  return c->mpibuf[mode].recvsize;
}

/* ------------------------------------------------------------------------------------------ */
/* moments (N4) and the diagnostic packers (N3)                                               */
/* ------------------------------------------------------------------------------------------ */
double* nixo_chunk_um(nixo_chunk* c)
{
  return c->um;
}

/* The reference ships the scatter (append_moment3d<Order>, primitives.hpp:896-930: um(iz0+jz, iy0+jy, ix0+jx,
 * is, k) += moment[jz][jy][jx][k]) but not the loop that fills moment[][][][14] -- it is the downstream
 * application's.  The composition used here (and by the GPU kernel), per particle of mass m, u = gamma v:
 *   weights  shape_mc<O> about the integer node, the stencil of the gather (base i - O/2 + Lb);
 *   moment[jz][jy][jx][k] = ((wz wy) wx) mom[k],   mom = { m;  m u_i/gamma (x,y,z);  m gamma c^2;  m u_i c (x,y,z);
 *                                                         m u_i u_j / gamma for (xx, yy, zz, xy, yz, zx) }          */
void nixo_chunk_deposit_moment(nixo_chunk* c, double cc)
{
  const int    order = c->g.order, is_odd = order % 2, half = order / 2, n1 = order + 1;
  const size_t ncell = (size_t)c->M[0] * c->M[1] * c->M[2];
  const int    my = c->M[1], mx = c->M[2], ns = c->ns;
  memset(c->um, 0, sizeof(double) * ncell * (size_t)ns * 14);
  const double del[3] = {c->g.del[0], c->g.del[1], c->g.del[2]};
  const double rc = 1 / cc;
  for (int is = 0; is < ns; is++) {
    particle_t* p = &c->up[is];
    for (int ip = 0; ip < p->Np; ip++) {
      const double* xu = &p->xu[(size_t)ip * NC];
      const double  pos[3] = {xu[2], xu[1], xu[0]}; /* z, y, x */
      double        w[3][MAXO + 2];
      int           i0[3];
      for (int a = 0; a < 3; a++) {
        const double lo  = c->lim[a][0];
        const double rdx = 1 / del[a];
        int          i   = nixo_digitize(pos[a], lo - 0.5 * del[a] * is_odd, rdx) - is_odd;
        nixo_shape_mc(order, pos[a], (lo + 0.5 * del[a]) + i * del[a], rdx, w[a]);
        i0[a] = i - half + c->Lb[a];
      }
      const double ux = xu[3], uy = xu[4], uz = xu[5], m = p->m;
      const double gam = nixo_lorentz_factor(ux, uy, uz, rc);
      const double mom[14] = {m,           m * ux / gam, m * uy / gam, m * uz / gam, m * gam * cc * cc,
                              m * ux * cc, m * uy * cc,  m * uz * cc,  m * ux * ux / gam, m * uy * uy / gam,
                              m * uz * uz / gam, m * ux * uy / gam, m * uy * uz / gam, m * uz * ux / gam};
      for (int jz = 0; jz < n1; jz++)
        for (int jy = 0; jy < n1; jy++)
          for (int jx = 0; jx < n1; jx++) {
            const double ww  = (w[0][jz] * w[1][jy]) * w[2][jx];
            double*      dst = &c->um[(((((size_t)(i0[0] + jz)) * my + (i0[1] + jy)) * mx + (i0[2] + jx)) * ns + is) * 14];
            for (int k = 0; k < 14; k++) dst[k] += ww * mom[k];
          }
    }
  }
}

/* XtensorPacker3D::decimate_size, xtensor_packer3d.hpp:232-241 */
static int decimate_size(int lb, int ub, int decimate)
{
  int size = ub - lb + 1;
  return (size <= decimate) ? 1 : size / decimate;
}

/* XtensorPacker3D::decimate_field (xtensor_packer3d.hpp:185-230): block averages of the interior of
 * x[Mz][My][Mx][nc], accumulated block offset by block offset (that order fixes the rounding) */
static int decimate_array(const nixo_chunk* c, const double* x, int nc, int decimate, double* out)
{
  const int sz = decimate_size(c->Lb[0], c->Ub[0], decimate), sy = decimate_size(c->Lb[1], c->Ub[1], decimate),
            sx = decimate_size(c->Lb[2], c->Ub[2], decimate);
  const int n = sz * sy * sx * nc;
  if (!out) return n;
  const int    bz = (c->Ub[0] - c->Lb[0] + 1) / sz, by = (c->Ub[1] - c->Lb[1] + 1) / sy, bx = (c->Ub[2] - c->Lb[2] + 1) / sx;
  const double factor = 1.0 / (bz * by * bx);
  memset(out, 0, sizeof(double) * (size_t)n);
  for (int kz = 0; kz < bz; kz++)
    for (int ky = 0; ky < by; ky++)
      for (int kx = 0; kx < bx; kx++)
        for (int jz = 0; jz < sz; jz++)
          for (int jy = 0; jy < sy; jy++)
            for (int jx = 0; jx < sx; jx++) {
              const int     iz = c->Lb[0] + jz * bz + kz, iy = c->Lb[1] + jy * by + ky, ix = c->Lb[2] + jx * bx + kx;
              const double* src = &x[(((size_t)iz * c->M[1] + iy) * c->M[2] + ix) * nc];
              double*       dst = &out[(((size_t)jz * sy + jy) * sx + jx) * nc];
              for (int ic = 0; ic < nc; ic++) dst[ic] += factor * src[ic];
            }
  return n;
}

/* XtensorPacker3D::pack_field (xtensor_packer3d.hpp:62-82): colocate E/B at the cell centres
 * (colocate_field_3d, :279-302), then decimate */
int nixo_chunk_pack_field(nixo_chunk* c, int decimate, double* out)
{
  if (!out) return decimate_array(c, c->uf, 6, decimate, NULL);
  const size_t ncell = (size_t)c->M[0] * c->M[1] * c->M[2];
  const int    my = c->M[1], mx = c->M[2];
  double*      y = (double*)calloc(ncell * 6, sizeof(double));
#define X6(z, yy, x, k) c->uf[((((size_t)(z)) * my + (yy)) * mx + (x)) * 6 + (k)]
  for (int iz = c->Lb[0]; iz <= c->Ub[0]; iz++)
    for (int iy = c->Lb[1]; iy <= c->Ub[1]; iy++)
      for (int ix = c->Lb[2]; ix <= c->Ub[2]; ix++) {
        double* d = &y[(((size_t)iz * my + iy) * mx + ix) * 6];
        d[0] = 0.50 * (X6(iz, iy, ix, 0) + X6(iz, iy, ix + 1, 0));
        d[1] = 0.50 * (X6(iz, iy, ix, 1) + X6(iz, iy + 1, ix, 1));
        d[2] = 0.50 * (X6(iz, iy, ix, 2) + X6(iz + 1, iy, ix, 2));
        d[3] = 0.25 * (X6(iz, iy, ix, 3) + X6(iz + 1, iy + 1, ix, 3) + X6(iz, iy + 1, ix, 3) + X6(iz + 1, iy, ix, 3));
        d[4] = 0.25 * (X6(iz, iy, ix, 4) + X6(iz + 1, iy, ix + 1, 4) + X6(iz + 1, iy, ix, 4) + X6(iz, iy, ix + 1, 4));
        d[5] = 0.25 * (X6(iz, iy, ix, 5) + X6(iz, iy + 1, ix + 1, 5) + X6(iz, iy, ix + 1, 5) + X6(iz, iy + 1, ix, 5));
      }
#undef X6
  int n = decimate_array(c, y, 6, decimate, out);
  free(y);
  return n;
}

/* XtensorPacker3D::pack_moment (xtensor_packer3d.hpp:84-104): decimate only.  which: 0 = uj (4), 1 = um (ns*14) */
int nixo_chunk_pack_moment(nixo_chunk* c, int which, int decimate, double* out)
{
  return which == 0 ? decimate_array(c, c->uj, 4, decimate, out) : decimate_array(c, c->um, c->ns * 14, decimate, out);
}

/* XtensorPacker3D::pack_tracer (xtensor_packer3d.hpp:122-140): the particles whose 64-bit id is negative,
 * in container order */
int nixo_chunk_pack_tracer(nixo_chunk* c, int is, double* out)
{
  particle_t* p = &c->up[is];
  int         n = 0;
  for (int ip = 0; ip < p->Np; ip++) {
    int64_t id;
    memcpy(&id, &p->xu[(size_t)ip * NC + 6], sizeof(id));
    if (id < 0) {
      if (out) memcpy(&out[(size_t)n * NC], &p->xu[(size_t)ip * NC], sizeof(double) * NC);
      n++;
    }
  }
  return n;
}
