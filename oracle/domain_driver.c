/* domain_driver.c -- multi-chunk driver with in-process "loop-back MPI" for the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Written once against oracle/nix_oracle.h and linked into BOTH the
 * plain-C restatement (liboracle) and the reference-backed library (oracle/_ref).
 *
 * The reference moves halo buffers with MPI_Isend/Irecv (chunk.hpp:507-567) even between chunks of
 * the same rank.  Its own halo tests replace MPI by a memcpy from a send slot to a recv slot
 * (unittest/test_xtensor_halo3d.cpp:85-93); this driver does the same for every chunk pair:
 *
 *   chunk A sends slot d to its neighbour in direction d (chunk.hpp:532-542); the neighbour B
 *   receives it in its slot -d, because from B the sender lies in direction -d
 *   (chunk.hpp:546-554: comm(1-dirz,1-diry,1-dirx), tag = B's id).  Hence
 *       B.recv[e]  <-  neighbour(B, e).send[opposite(e)].
 *
 * Threading: OpenMP parallel-for over chunks inside each phase = the reference's threading model
 * (nix.hpp:39-49: chunks driven from OpenMP threads).
 */
#include "nix_oracle.h"

#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

struct nixo_domain {
  int          cdims[3];
  int          nchunk;
  int          ns;
  int          order;
  int          dims[3], nb;
  double       del[3];
  int*         coord;    /* [nchunk][3] */
  int*         grid2id;  /* [Cz][Cy][Cx] -> chunk index */
  int*         nbr;      /* [nchunk][27] */
  nixo_chunk** chunk;
  int*         sendsize; /* [nchunk][27] scratch */
  int*         sendaddr; /* [nchunk][27] scratch */
};

static int wrap(int a, int n)
{
  return ((a % n) + n) % n;
}

nixo_domain* nixo_domain_create(const int* cdims, const int* dims, int nb, int order,
                                const double* del, int ns, const double* q, const double* m,
                                const int* coord, const int* np_required)
{
  nixo_domain* d = (nixo_domain*)calloc(1, sizeof(nixo_domain));
  d->nchunk      = cdims[0] * cdims[1] * cdims[2];
  d->ns          = ns;
  d->order       = order;
  memcpy(d->cdims, cdims, 3 * sizeof(int));
  memcpy(d->dims, dims, 3 * sizeof(int));
  memcpy(d->del, del, 3 * sizeof(double));
  d->nb = nb;
  d->coord    = (int*)malloc(sizeof(int) * 3 * d->nchunk);
  d->grid2id  = (int*)malloc(sizeof(int) * d->nchunk);
  d->nbr      = (int*)malloc(sizeof(int) * 27 * d->nchunk);
  d->chunk    = (nixo_chunk**)malloc(sizeof(nixo_chunk*) * d->nchunk);
  d->sendsize = (int*)malloc(sizeof(int) * 27 * d->nchunk);
  d->sendaddr = (int*)malloc(sizeof(int) * 27 * d->nchunk);
  memcpy(d->coord, coord, sizeof(int) * 3 * d->nchunk);

  for (int k = 0; k < d->nchunk; k++) {
    const int* c = &coord[3 * k];
    d->grid2id[(c[0] * cdims[1] + c[1]) * cdims[2] + c[2]] = k;
  }
  for (int k = 0; k < d->nchunk; k++) {
    const int* c = &coord[3 * k];
    for (int iz = 0; iz < 3; iz++)
      for (int iy = 0; iy < 3; iy++)
        for (int ix = 0; ix < 3; ix++) {
          int nz = wrap(c[0] + iz - 1, cdims[0]);
          int ny = wrap(c[1] + iy - 1, cdims[1]);
          int nx = wrap(c[2] + ix - 1, cdims[2]);
          d->nbr[27 * k + 9 * iz + 3 * iy + ix] = d->grid2id[(nz * cdims[1] + ny) * cdims[2] + nx];
        }
    nixo_geom g;
    for (int a = 0; a < 3; a++) {
      g.dims[a]   = dims[a];
      g.offset[a] = c[a] * dims[a];
      g.gdims[a]  = cdims[a] * dims[a];
      g.del[a]    = del[a];
    }
    g.nb        = nb;
    g.order     = order;
    d->chunk[k] = nixo_chunk_create(&g, ns, &np_required[ns * k], q, m);
  }
  return d;
}

void nixo_domain_destroy(nixo_domain* d)
{
  if (!d)
    return;
  for (int k = 0; k < d->nchunk; k++)
    nixo_chunk_destroy(d->chunk[k]);
  free(d->coord);
  free(d->grid2id);
  free(d->nbr);
  free(d->chunk);
  free(d->sendsize);
  free(d->sendaddr);
  free(d);
}

int nixo_domain_nchunk(nixo_domain* d)
{
  return d->nchunk;
}

nixo_chunk* nixo_domain_chunk(nixo_domain* d, int k)
{
  return d->chunk[k];
}

int nixo_domain_neighbor(nixo_domain* d, int k, int iz, int iy, int ix)
{
  return d->nbr[27 * k + 9 * iz + 3 * iy + ix];
}

void nixo_domain_clear_current(nixo_domain* d)
{
  nixo_geom g; /* size only */
  (void)g;
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < d->nchunk; k++) {
    /* uj size = Mz*My*Mx*4; recover from the current buffer layout: centre slot excluded, so
     * keep it simple and ask the chunk for Ng (= Mz*My*Mx, particle.hpp:99-106). */
    int     ng = nixo_particle_ng(d->chunk[k], 0);
    double* uj = nixo_chunk_uj(d->chunk[k]);
    memset(uj, 0, sizeof(double) * 4 * (size_t)ng);
  }
}

void nixo_domain_push_deposit(nixo_domain* d, double delt, double cc, int simd)
{
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < d->nchunk; k++) {
    nixo_chunk_push_deposit(d->chunk[k], delt, cc, simd);
  }
}

void nixo_domain_exchange(nixo_domain* d, int mode)
{
  /* particle mode: count(0, Np-1, reset=true) precedes pack (test_xtensor_halo3d.cpp:186) */
  if (mode == NIXO_MODE_PARTICLE) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < d->nchunk; k++) {
      for (int is = 0; is < d->ns; is++) {
        int np = nixo_particle_np(d->chunk[k], is);
        nixo_particle_count(d->chunk[k], is, 0, np - 1, 1, d->order);
      }
    }
  }

  /* pack_bc_exchange on every chunk; remember the send layout (the reference overwrites
   * bufsize/bufaddr with the receive layout in probe_bc_exchange, chunk.cpp:318-319) */
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < d->nchunk; k++) {
    nixo_chunk_halo_pack(d->chunk[k], mode);
    for (int s = 0; s < 27; s++) {
      d->sendsize[27 * k + s] = nixo_chunk_bufsize(d->chunk[k], mode, s / 9, (s / 3) % 3, s % 3);
      d->sendaddr[27 * k + s] = nixo_chunk_bufaddr(d->chunk[k], mode, s / 9, (s / 3) % 3, s % 3);
    }
    d->sendsize[27 * k + 13] = 0;
  }

  /* loop-back "MPI": B.recv[e] <- neighbour(B,e).send[26-e] */
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < d->nchunk; k++) {
    int sizes[27];
    for (int e = 0; e < 27; e++) {
      int n    = d->nbr[27 * k + e];
      sizes[e] = (e == 13) ? 0 : d->sendsize[27 * n + (26 - e)];
    }
    nixo_chunk_set_recv_sizes(d->chunk[k], mode, sizes);
    uint8_t* recv = nixo_chunk_recvbuf(d->chunk[k], mode);
    for (int e = 0; e < 27; e++) {
      if (e == 13)
        continue;
      int      n    = d->nbr[27 * k + e];
      uint8_t* send = nixo_chunk_sendbuf(d->chunk[n], mode);
      int      addr = nixo_chunk_bufaddr(d->chunk[k], mode, e / 9, (e / 3) % 3, e % 3);
      memcpy(recv + addr, send + d->sendaddr[27 * n + (26 - e)], (size_t)sizes[e]);
    }
  }

  /* unpack_bc_exchange on every chunk (particle mode: append + wrap + count + sort) */
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < d->nchunk; k++) {
    nixo_chunk_halo_unpack(d->chunk[k], mode);
  }
}

void nixo_domain_sort_only(nixo_domain* d)
{
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < d->nchunk; k++) {
    for (int is = 0; is < d->ns; is++) {
      int np = nixo_particle_np(d->chunk[k], is);
      nixo_particle_count(d->chunk[k], is, 0, np - 1, 1, d->order);
      nixo_particle_sort(d->chunk[k], is);
    }
  }
}

void nixo_domain_step(nixo_domain* d, double delt, double cc, int simd)
{
  nixo_domain_clear_current(d);
  nixo_domain_push_deposit(d, delt, cc, simd);
  nixo_domain_exchange(d, NIXO_MODE_CURRENT);
  /* [field solver would run here -- downstream of nix, not on this path] */
  nixo_domain_exchange(d, NIXO_MODE_FIELD);
  nixo_domain_exchange(d, NIXO_MODE_PARTICLE);
}

/* ---- the same step with the Yee field update on the grid (oracle/field_solver.c; NOT in the reference
 *      tree -- parity unpinned by the reference).  Time-centred leapfrog:
 *        push with E^n, B^n -> J^{n+1/2};  B^n -> B^{n+1/2} (on the cells one ghost layer out as well,
 *        from the ghost E of the last exchange, so that E needs no exchange of B);  E^n -> E^{n+1};
 *        E/B halo;  B^{n+1/2} -> B^{n+1};  E/B halo;  particle migration + sort ---- */
void nixo_domain_push_bfd(nixo_domain* d, double delt, double cc, int ext)
{
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < d->nchunk; k++)
    nixo_fdtd_push_bfd(nixo_chunk_uf(d->chunk[k]), d->dims, d->nb, d->del, cc, delt, ext);
}

void nixo_domain_push_efd(nixo_domain* d, double delt, double cc, double cfj)
{
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < d->nchunk; k++)
    nixo_fdtd_push_efd(nixo_chunk_uf(d->chunk[k]), nixo_chunk_uj(d->chunk[k]), d->dims, d->nb, d->del, cc, delt, cfj);
}

void nixo_domain_step_em(nixo_domain* d, double delt, double cc, double cfj, int simd)
{
  nixo_domain_clear_current(d);
  nixo_domain_push_deposit(d, delt, cc, simd);
  nixo_domain_exchange(d, NIXO_MODE_CURRENT);
  nixo_domain_push_bfd(d, 0.5 * delt, cc, 1);
  nixo_domain_push_efd(d, delt, cc, cfj);
  nixo_domain_exchange(d, NIXO_MODE_FIELD);
  nixo_domain_push_bfd(d, 0.5 * delt, cc, 0);
  nixo_domain_exchange(d, NIXO_MODE_FIELD);
  nixo_domain_exchange(d, NIXO_MODE_PARTICLE);
}

void nixo_domain_deposit_moment(nixo_domain* d, double cc)
{
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < d->nchunk; k++)
    nixo_chunk_deposit_moment(d->chunk[k], cc);
  nixo_domain_exchange(d, NIXO_MODE_MOMENT);
}

int64_t nixo_domain_total_particles(nixo_domain* d)
{
  int64_t n = 0;
  for (int k = 0; k < d->nchunk; k++)
    for (int is = 0; is < d->ns; is++)
      n += nixo_particle_np(d->chunk[k], is);
  return n;
}

void nixo_set_num_threads(int n)
{
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int nixo_get_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
