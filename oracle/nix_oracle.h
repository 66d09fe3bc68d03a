/* nix_oracle.h -- C API of the CPU ORACLE for the per-chunk PIC step of amanotk/nix.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * The SAME API is implemented twice:
 *   oracle/nix_oracle.c        plain-C restatement of the reference algorithm (travels everywhere)
 *   oracle/ref/ref_driver.cpp  thin driver that calls the reference's OWN templates/classes from
 *                              /root/reference (built into oracle/_ref/, never copied)
 * and tests/test_oracle_vs_ref.py requires the two to agree bit for bit.
 * oracle/domain_driver.c (multi-chunk loop-back exchange) is written once against this API and is
 * linked into both libraries.
 *
 * Index convention everywhere: axis order (z, y, x); 27 directions are indexed [iz][iy][ix] with
 * 0/1/2 = -/centre/+  (reference chunk.hpp:11-41, chunk.cpp:171-207).
 */
#ifndef NIX_ORACLE_H
#define NIX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NIXO_MODE_FIELD 0    /* XtensorHaloField3D    xtensor_halo3d.hpp:19-71   */
#define NIXO_MODE_CURRENT 1  /* XtensorHaloCurrent3D  xtensor_halo3d.hpp:77-129  */
#define NIXO_MODE_PARTICLE 2 /* XtensorHaloParticle3D xtensor_halo3d.hpp:251-557 */
#define NIXO_MODE_MOMENT 3   /* XtensorHaloMoment3D   xtensor_halo3d.hpp:135-187 */

/* Geometry of one chunk: exactly the inputs of Chunk::Chunk / set_boundary_margin /
 * set_global_context / set_coordinate (chunk.cpp:6-16,118-247). */
typedef struct {
  int    dims[3];   /* cells in the chunk (Nz, Ny, Nx) */
  int    nb;        /* boundary margin (ghost width) */
  int    order;     /* shape-function order 1..3 */
  int    offset[3]; /* global cell offset of the chunk (z, y, x) */
  int    gdims[3];  /* global number of cells (z, y, x) */
  double del[3];    /* grid spacing (delz, dely, delx) */
} nixo_geom;

typedef struct nixo_chunk  nixo_chunk;
typedef struct nixo_domain nixo_domain;

const char* nixo_impl_name(void); /* "port" or "reference" */
int         nixo_simd_lanes(void);

/* ---- chunk ---- */
nixo_chunk* nixo_chunk_create(const nixo_geom* g, int ns, const int* np_required, const double* q,
                              const double* m);
void        nixo_chunk_destroy(nixo_chunk* c);
double*     nixo_chunk_uf(nixo_chunk* c); /* [Mz][My][Mx][6]  Ex Ey Ez Bx By Bz */
double*     nixo_chunk_uj(nixo_chunk* c); /* [Mz][My][Mx][4]  rho Jx Jy Jz     */
void        nixo_chunk_set_nb_valid(nixo_chunk* c, int iz, int iy, int ix, int valid);

/* ---- particle container of species `is` (XtensorParticle, xtensor_particle.hpp) ---- */
int      nixo_particle_ng(nixo_chunk* c, int is);
int      nixo_particle_np(nixo_chunk* c, int is);
void     nixo_particle_set_np(nixo_chunk* c, int is, int np);
int      nixo_particle_np_total(nixo_chunk* c, int is);
double*  nixo_particle_xu(nixo_chunk* c, int is);     /* [Np_total][7] */
double*  nixo_particle_xv(nixo_chunk* c, int is);     /* [Np_total][7] */
int32_t* nixo_particle_gindex(nixo_chunk* c, int is); /* [Np_total]    */
int32_t* nixo_particle_pindex(nixo_chunk* c, int is); /* [Ng+1]        */
int32_t* nixo_particle_pcount(nixo_chunk* c, int is); /* [Ng+1][8]     */
void     nixo_particle_resize(nixo_chunk* c, int is, int np_required);
void     nixo_particle_count(nixo_chunk* c, int is, int lbp, int ubp, int reset, int order);
void     nixo_particle_sort(nixo_chunk* c, int is);
void     nixo_particle_set_boundary_periodic(nixo_chunk* c, int is, int lbp, int ubp);

/* ---- numerical primitives (scalar instantiations), for known-answer tests ---- */
int    nixo_digitize(double x, double xmin, double rdx);                 /* primitives.hpp:46-58  */
void   nixo_shape_mc(int order, double x, double X, double rdx, double* s); /* order 1..4  :257-331,519-532 */
void   nixo_shape_wt(int order, double x, double X, double rdx, double dt, double rdt, double* s); /* :333-495,555-570 */
void   nixo_push_boris(double* u, const double* eb, double cc);          /* primitives.hpp:165-189 */
void   nixo_push_vay(double* u, const double* eb, double cc);            /* primitives.hpp:193-224 */
void   nixo_push_higuera_cary(double* u, const double* eb, double cc);   /* primitives.hpp:227-253 */
void   nixo_set_pusher(int pusher); /* composed step: 0 Boris, 1 Vay, 2 Higuera-Cary */
int    nixo_get_pusher(void);
double nixo_lorentz_factor(double ux, double uy, double uz, double rc);  /* primitives.hpp:158-161 */
/* esirkepov::deposit3d<order> on ss[2][3][order+3] -> cur[(order+3)^3][4] (esirkepov.hpp:326-340) */
void nixo_deposit3d(int order, double dxdt, double dydt, double dzdt, double qs, double* ss,
                    double* cur);
/* interp::shift_weights<Order> (interp.hpp:149-160), ww[order+2]; esirkepov::shift_weights<3, Order>
 * (esirkepov.hpp:241-258), ss[3][order+3], shift[3] */
void nixo_interp_shift_weights(int order, int shift, double* ww);
void nixo_esirkepov_shift_weights(int order, const int* shift, double* ss);
/* append_current3d<order> (primitives.hpp:778-834) on uj[mz][my][mx][4] with cur[(order+3)^3][4];
 * append_moment3d<order> (primitives.hpp:896-930) on um[mz][my][mx][ns][14] with mom[(order+1)^3][14]; mz = the
 * highest plane touched + 1 */
void nixo_append_current3d(int order, double* uj, int my, int mx, int iz0, int iy0, int ix0, const double* cur);
void nixo_append_moment3d(int order, double* um, int my, int mx, int ns, int iz0, int iy0, int ix0, int is,
                          const double* mom);
/* interp::interp3d<order> scalar (interp.hpp:95-113,217-230) on a [Mz][My][Mx][6] array */
double nixo_interp3d(int order, const double* eb, int my, int mx, int iz0, int iy0, int ix0, int ik,
                     const double* wz, const double* wy, const double* wx, double dt);

/* ---- composed per-chunk step, part (a)+(b) of SURVEY 3.2: gather + Boris push + position update
 *      + Esirkepov deposit into uj, for every species.  simd=0: scalar templates; simd=1 (reference
 *      build only): xsimd batches over cell-sorted particles (the sorted code paths). ---- */
void nixo_chunk_push_deposit(nixo_chunk* c, double delt, double cc, int simd);

/* ---- halo engine: Chunk::pack_bc_exchange / unpack_bc_exchange with the three Halo classes ---- */
void     nixo_chunk_halo_pack(nixo_chunk* c, int mode);
void     nixo_chunk_halo_unpack(nixo_chunk* c, int mode);
int      nixo_chunk_bufsize(nixo_chunk* c, int mode, int iz, int iy, int ix);
int      nixo_chunk_bufaddr(nixo_chunk* c, int mode, int iz, int iy, int ix);
uint8_t* nixo_chunk_sendbuf(nixo_chunk* c, int mode); /* whole send buffer */
int      nixo_chunk_sendbuf_size(nixo_chunk* c, int mode);
/* emulate probe_bc_exchange (chunk.cpp:310-395): set per-direction sizes, lay out + resize recvbuf */
void     nixo_chunk_set_recv_sizes(nixo_chunk* c, int mode, const int* bufsize27);
uint8_t* nixo_chunk_recvbuf(nixo_chunk* c, int mode);
int      nixo_chunk_recvbuf_size(nixo_chunk* c, int mode);

/* ---- multi-chunk domain with loop-back exchange (oracle/domain_driver.c) ---- */
/* cdims = number of chunks (Cz, Cy, Cx); every chunk has `dims` cells; periodic in all directions.
 * Chunk k has coordinates coord[3*k..3*k+2] (cz, cy, cx) -- the caller supplies the SFC order. */
nixo_domain* nixo_domain_create(const int* cdims, const int* dims, int nb, int order,
                                const double* del, int ns, const double* q, const double* m,
                                const int* coord, const int* np_required /* [nchunk][ns] */);
void         nixo_domain_destroy(nixo_domain* d);
int          nixo_domain_nchunk(nixo_domain* d);
nixo_chunk*  nixo_domain_chunk(nixo_domain* d, int k);
int          nixo_domain_neighbor(nixo_domain* d, int k, int iz, int iy, int ix);
void         nixo_domain_clear_current(nixo_domain* d);
void         nixo_domain_push_deposit(nixo_domain* d, double delt, double cc, int simd);
void         nixo_domain_exchange(nixo_domain* d, int mode); /* pack -> loop-back -> unpack */
void         nixo_domain_sort_only(nixo_domain* d);          /* count(reset) + sort, no exchange */
/* one full step: clear J, push+deposit, J halo, E/B halo, particle migration (count+pack+unpack+sort) */
void    nixo_domain_step(nixo_domain* d, double delt, double cc, int simd);
int64_t nixo_domain_total_particles(nixo_domain* d);

/* ---- moments (append_moment3d primitives.hpp:896-930, XtensorHaloMoment3D) and the diagnostic packers
 *      (XtensorPacker3D, xtensor_packer3d.hpp:62-140); `out` may be NULL (query: returns the element count) ---- */
double* nixo_chunk_um(nixo_chunk* c); /* [Mz][My][Mx][ns][14] */
void    nixo_chunk_deposit_moment(nixo_chunk* c, double cc);
int     nixo_chunk_pack_field(nixo_chunk* c, int decimate, double* out);
int     nixo_chunk_pack_moment(nixo_chunk* c, int which /* 0: uj, 1: um */, int decimate, double* out);
int     nixo_chunk_pack_tracer(nixo_chunk* c, int is, double* out);
void    nixo_domain_deposit_moment(nixo_domain* d, double cc); /* every chunk + the moment halo */

/* ---- Yee FDTD field update (oracle/field_solver.c).  NOT part of the reference tree: parity unpinned by
 *      the reference, pinned by known answers (tests/test_field_solver.py).  uf [Mz][My][Mx][6],
 *      uj [Mz][My][Mx][4]; ext = extra cell layers around the interior that are updated too ---- */
void nixo_fdtd_push_bfd(double* uf, const int* dims, int nb, const double* del, double cc, double delt, int ext);
void nixo_fdtd_push_efd(double* uf, const double* uj, const int* dims, int nb, const double* del, double cc,
                        double delt, double cfj);
void nixo_fdtd_energy(const double* uf, const int* dims, int nb, double* e2b2);
void nixo_domain_push_bfd(nixo_domain* d, double delt, double cc, int ext);
void nixo_domain_push_efd(nixo_domain* d, double delt, double cc, double cfj);
/* clear J, push+deposit, J halo, B half step, E step, E/B halo, B half step, E/B halo, migration */
void nixo_domain_step_em(nixo_domain* d, double delt, double cc, double cfj, int simd);
void    nixo_set_num_threads(int n);
int     nixo_get_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
