/* nixb200.h -- C ABI of the B200-native per-chunk PIC step for amanotk/nix.
 *
 * This is the drop-in boundary: a C-ABI shared library (libnixb200.so) whose entry points are what
 * a `GpuChunk : nix::Chunk` / `GpuApplication : nix::Application` pair would bind (see
 * INTEGRATION.md).  The reference has no FFI of its own -- its extension API is C++ inheritance
 * (application.hpp:45-48,343-346; chunk.hpp:139-192,399-432) -- so every entry point cites the
 * reference interface it stands behind.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; nixb200_last_error() gives the text
 *     (the reference logs with ERROR and carries on; no exceptions cross this boundary)
 *   - axis order is (z, y, x) everywhere; the 27 directions are indexed 9*iz + 3*iy + ix with
 *     0/1/2 = -/centre/+ (chunk.hpp:11-41)
 *   - particles cross the boundary in the reference's AoS layout xu[np][7] =
 *     (x, y, z, ux, uy, uz, id-bits) (xtensor_particle.hpp:15, particle.hpp:18); on the device they
 *     are SoA
 *   - fields cross as uf[Mz][My][Mx][6] (Ex Ey Ez Bx By Bz) and uj[Mz][My][Mx][4] (rho Jx Jy Jz),
 *     M = N + 2*nb, exactly the reference's xtensor arrays
 *   - all work is enqueued on the domain's stream (nixb200_domain_set_stream); calls that return
 *     data to the host synchronise that stream
 *   - there is NO CPU fallback: every entry point fails if no sm_100 device is present
 */
#ifndef NIXB200_H
#define NIXB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NIXB200_MODE_FIELD 0    /* XtensorHaloField3D    xtensor_halo3d.hpp:19-71   */
#define NIXB200_MODE_CURRENT 1  /* XtensorHaloCurrent3D  xtensor_halo3d.hpp:77-129  */
#define NIXB200_MODE_PARTICLE 2 /* XtensorHaloParticle3D xtensor_halo3d.hpp:251-557 */
#define NIXB200_MODE_MOMENT 3   /* XtensorHaloMoment3D   xtensor_halo3d.hpp:135-187 */

#define NIXB200_FIELD_UF 0
#define NIXB200_FIELD_UJ 1

/* error bits accumulated on the device, see nixb200_domain_check() */
#define NIXB200_ERR_UNSORTED 1 /* push found a particle outside the cell it is binned in */
#define NIXB200_ERR_CFL 2      /* a particle moved more than one cell (c*dt > dx)          */
#define NIXB200_ERR_CAPACITY 4 /* a particle / message buffer overflowed: particles were lost; the
                                  domain refuses further steps until particles are set again     */

/* momentum update of the push: the reference's three interchangeable primitives */
#define NIXB200_PUSH_BORIS 0        /* push_boris         primitives.hpp:165-189 */
#define NIXB200_PUSH_VAY 1          /* push_vay           primitives.hpp:193-224 */
#define NIXB200_PUSH_HIGUERA_CARY 2 /* push_higuera_cary  primitives.hpp:227-253 */

typedef struct nixb200_domain nixb200_domain;

/* A domain = the chunks of ONE rank (one GPU): a contiguous range of chunk ids along the
 * space-filling curve, exactly ChunkMap::get_rank semantics (chunkmap.cpp:156-164).  The host keeps
 * ChunkMap / Balancer / SFC; it hands us the id -> (cz,cy,cx) table and the rank boundaries. */
typedef struct {
  int    cdims[3];   /* chunks per axis of the global periodic box (Cz, Cy, Cx)  cfgparser.hpp:187  */
  int    dims[3];    /* cells per chunk (Nz, Ny, Nx)                             chunk.cpp:6-16     */
  int    nb;         /* boundary margin                                          chunk.hpp:260-264  */
  int    order;      /* shape order 1..3                                         primitives.hpp:519 */
  int    ns;         /* number of species                                                           */
  double del[3];     /* delz, dely, delx                                         chunk.cpp:210-237  */
  double cc;         /* speed of light                                                              */
  int    id_begin;   /* first chunk id owned by this rank   (boundary[rank])     chunkmap.cpp:156   */
  int    id_end;     /* one past the last chunk id owned    (boundary[rank+1])                      */
  int    device;     /* CUDA device ordinal                                                         */
  int    strict_fp;  /* 1: push arithmetic without FMA contraction, bit-identical to the reference's
                        scalar templates; 0: contracted (<=1e-12 relative)                           */
  double capacity_factor; /* initial particle storage = factor * initial count (>=1; 0 -> 1.25); the
                             stores regrow on their own while a run drifts (see nixb200_domain_reserve) */
  int    pusher;     /* NIXB200_PUSH_BORIS / _VAY / _HIGUERA_CARY               primitives.hpp:165-253 */
  int    fp32;       /* 0: fp64 on the device (the reference's real type, nix.hpp:76-78); 1: fp32 mode -- particles,
                        E/B and J are kept as floats on the device (particle positions relative to their chunk's
                        origin), all kernels compute in fp32; this C ABI stays fp64 (converted at the boundary).
                        No reference exists for it: results agree with the fp64 path to ~1e-6 (tests: 1e-5);
                        strict_fp is ignored; the MpiBuffer-layout halo calls and the full-array overlapped
                        transfers are fp64 only */
} nixb200_domain_desc;

const char* nixb200_last_error(void);
const char* nixb200_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t nixb200_launch_count(void);

/* coord: [ncid][3] chunk id -> (cz,cy,cx) for ALL chunk ids of the box (ChunkMap::get_coordinate,
 * chunkmap.cpp:176-191); q, m: [ns] charge and mass per species (particle.hpp:24-25). */
int nixb200_domain_create(const nixb200_domain_desc* desc, const int* coord, const double* q,
                          const double* m, nixb200_domain** out);
int nixb200_domain_destroy(nixb200_domain* d);
int nixb200_domain_set_stream(nixb200_domain* d, void* cuda_stream);
int nixb200_domain_synchronize(nixb200_domain* d);
/* returns and clears the device error bits (NIXB200_ERR_*) */
int nixb200_domain_check(nixb200_domain* d, int* errbits);

/* ---- data in / out (backs Chunk::pack/unpack, chunk.cpp:18-116; XtensorParticle::pack/unpack,
 *      xtensor_particle.hpp:128-218; and the diagnostics download path) ---- */
/* local chunk index k = id - id_begin */
int nixb200_chunk_field_upload(nixb200_domain* d, int k, int which, const double* host);
int nixb200_chunk_field_download(nixb200_domain* d, int k, int which, double* host);
/* all local chunks at once, asynchronous on the domain's stream (host memory should be pinned);
 * call nixb200_domain_synchronize before reusing / reading the host buffer */
int nixb200_domain_field_upload_async(nixb200_domain* d, int which, const double* host);
int nixb200_domain_field_download_async(nixb200_domain* d, int which, double* host);
/* Overlapped variants: the copy runs on a second stream of the domain and is ordered by events against
 * the phases that touch the array (uf: push_deposit / exchange_field, uj: clear_current / push_deposit /
 * exchange_current).  A download of uj issued after exchange_current and an upload of the next step's uf
 * issued after exchange_field run while the main stream migrates and sorts particles (the slot where
 * Chunk-side code of a nix application solves the fields, application.cpp:63-69).  `host` must be
 * page-locked for the copy to be asynchronous; copy_synchronize waits for the copies only. */
int nixb200_domain_field_upload_overlapped(nixb200_domain* d, int which, const double* host);
int nixb200_domain_field_download_overlapped(nixb200_domain* d, int which, double* host);
/* The same for the INTERIOR cells only: host layout [chunk][Nz][Ny][Nx][6 or 4], no ghost cells (a host-side
 * solver needs the interior J -- the J halo has folded the ghosts in -- and produces the interior E/B, whose
 * ghosts nixb200_domain_exchange_field fills on the device).  About half the bytes at Nb = 2, N = 16. */
int nixb200_domain_interior_upload_overlapped(nixb200_domain* d, int which, const double* host);
int nixb200_domain_interior_download_overlapped(nixb200_domain* d, int which, double* host);
int nixb200_domain_copy_synchronize(nixb200_domain* d);
/* all chunks of one species at once: xu_aos = concatenation over local chunks, np_chunk[k] each (host or
 * device memory) */
int nixb200_domain_set_particles(nixb200_domain* d, int is, const double* xu_aos,
                                 const int64_t* np_chunk);
int nixb200_domain_get_np(nixb200_domain* d, int is, int64_t* np_chunk /* [nchunk] */);
int nixb200_chunk_get_particles(nixb200_domain* d, int k, int is, double* xu_aos, int64_t max_np,
                                int64_t* np);
/* pindex[Ng+1] / pcount[Ng+1][8] in the reference's layout (xtensor_particle.hpp:18-19): pcount is
 * the state count() leaves (before sort() turns it into cursors); pindex the state sort() leaves */
int nixb200_chunk_get_pindex(nixb200_domain* d, int k, int is, int32_t* pindex);
int nixb200_chunk_get_pcount(nixb200_domain* d, int k, int is, int32_t* pcount);

/* Particle storage.  The reference's containers grow on demand (XtensorParticle::resize called from
 * XtensorHaloParticle3D::pre_unpack, xtensor_halo3d.hpp:464-476).  Here each species has one store of `np`
 * particles for all chunks of the domain and migration buffers for `nmove` particles per step; after every
 * sort the device reports its fill level and the next push_deposit regrows the stores BEFORE they are full
 * (what moved in the last step must fit twice on top of what is there).  A jump that outruns this in a single
 * step raises NIXB200_ERR_CAPACITY -- never an out-of-bounds access -- and is avoided by reserving up front. */
int nixb200_domain_reserve(nixb200_domain* d, int is, int64_t np, int64_t nmove);
int nixb200_domain_get_capacity(nixb200_domain* d, int is, int64_t* np, int64_t* nmove);

/* ---- the per-step hot path, batched over every chunk of the domain ---- */
/* XtensorParticle::count + sort for every chunk/species (xtensor_particle.hpp:260-357); drops
 * out-of-bounds particles exactly as the reference does.  Needed once after set_particles. */
int nixb200_domain_sort(nixb200_domain* d);
int nixb200_domain_clear_current(nixb200_domain* d);
/* gather (interp3d) + push_boris + position update + Esirkepov deposit3d/append_current3d for every
 * species; also produces the per-cell counts of the NEW positions (count(0,Np-1,true,order)) */
int nixb200_domain_push_deposit(nixb200_domain* d, double delt);
/* set_boundary_{pack,begin,end,unpack}(mode) for all local chunk pairs (chunk.hpp:435-586) */
int nixb200_domain_exchange_current(nixb200_domain* d);
int nixb200_domain_exchange_field(nixb200_domain* d);
/* XtensorHaloParticle3D pack -> exchange -> unpack (append, periodic wrap, count, sort) */
int nixb200_domain_migrate_sort(nixb200_domain* d);
/* clear J, push+deposit, J halo, E/B halo, migrate+sort: one Application::push() worth of work */
int nixb200_domain_step(nixb200_domain* d, double delt);

/* ---- field solver on the device.  The reference has none -- Application::push() is an
 *      empty virtual (application.hpp:343-346), the Maxwell update belongs to the downstream application --
 *      but it fixes the Yee staggering of uf (xtensor_packer3d.hpp:279-302) and of J (esirkepov.hpp:177-237).
 *      With these calls E, B and J never leave the device between steps.
 *        push_bfd   B -= c dt curl E on the interior plus `ext` (< nb) ghost layers
 *        push_efd   E += c dt curl B - cfj dt J on the interior (cfj: 1 in Heaviside-Lorentz units, 4 pi in Gaussian)
 *        step_em    clear J, push+deposit, J halo, B half step (ext = 1: no exchange of B needed), E step,
 *                   E/B halo, B half step, E/B halo, migrate+sort  -- a time-centred leapfrog
 *        field_energy   [nchunk][2] = sum E^2, sum B^2 over the interior cells of every chunk (history diagnostic) ---- */
int nixb200_domain_push_bfd(nixb200_domain* d, double delt, int ext);
int nixb200_domain_push_efd(nixb200_domain* d, double delt, double cfj);
int nixb200_domain_step_em(nixb200_domain* d, double delt, double cfj);
int nixb200_domain_field_energy(nixb200_domain* d, double* host_e2b2);
/* the per-step history read-back without stalling the device: field energies [nchunk][2] and particle counts
 * [ns][nchunk] (int64) arrive in caller-provided PINNED host memory in stream order (either pointer may be NULL);
 * read them after nixb200_domain_synchronize or any later call that synchronises */
int nixb200_domain_history_async(nixb200_domain* d, double* host_e2b2, int64_t* host_np);
/* 1: push arithmetic without FMA contraction (bit-identical to the reference's scalar templates), 0: contracted */
int nixb200_domain_set_strict_fp(nixb200_domain* d, int on);

/* ---- diagnostics / output on the device (SURVEY.md 8f, N3 + N4): only the packed output crosses PCIe ----
 * moments: um[Mz][My][Mx][ns][14] per chunk = append_moment3d<Order> (primitives.hpp:896-930) over every
 *   particle, then XtensorHaloMoment3D (xtensor_halo3d.hpp:135-187) incl. neighbours on other ranks.  The
 *   14 moments per particle (the loop is the downstream application's; ours, as in the oracle):
 *   m; m u_i/gamma (x,y,z); m gamma c^2; m u_i c (x,y,z); m u_i u_j/gamma (xx,yy,zz,xy,yz,zx)
 * pack_field   XtensorPacker3D::pack_field  (xtensor_packer3d.hpp:62-82): E/B colocated at the cell centres,
 *              then block-averaged by `decimate`  -> host[sz][sy][sx][6]
 * pack_moment  XtensorPacker3D::pack_moment (:84-104): block averages of uj (which = 0, 4 values) or um
 *              (which = 1, ns*14 values)
 * pack_tracer  XtensorPacker3D::pack_tracer (:122-140): the particles of a chunk with a negative 64-bit id, in
 *              container order, AoS [n][7]
 * host == NULL queries the element count.  pack_field / pack_moment reproduce the reference's rounding. */
int nixb200_domain_deposit_moment(nixb200_domain* d);
int nixb200_chunk_moment_download(nixb200_domain* d, int k, double* host);
int nixb200_chunk_pack_field(nixb200_domain* d, int k, int decimate, double* host, int64_t* count);
int nixb200_chunk_pack_moment(nixb200_domain* d, int k, int which, int decimate, double* host, int64_t* count);
int nixb200_chunk_pack_tracer(nixb200_domain* d, int k, int is, double* host_aos, int64_t max_np, int64_t* np);
/* the shape functions the push does not use, as device primitives (bit-identical to the reference's scalar
 * templates): kind 0 = shape_mc<order> (order 1..4, primitives.hpp:257-331), 1 = shape_wt<order> (:333-495);
 * out[n][order+1] */
int nixb200_shape_eval(int device, int kind, int order, int n, const double* x, const double* X, double rdx, double dt,
                       double rdt, double* out);

/* ---- per-chunk halo buffers in the reference's MpiBuffer layout (chunk.cpp:257-286), for
 *      neighbours that live on another rank and for drop-in use behind Chunk::set_boundary_* ---- */
int nixb200_halo_layout(nixb200_domain* d, int mode, int* bufsize27, int* bufaddr27);
/* the same from the chunk shape alone (host logic, no device): dims = (Nz, Ny, Nx) */
int nixb200_halo_layout_dims(const int* dims3, int nb, int mode, int* bufsize27, int* bufaddr27);
int nixb200_chunk_halo_pack(nixb200_domain* d, int k, int mode, void* host_sendbuf);
int nixb200_chunk_halo_unpack(nixb200_domain* d, int k, int mode, const void* host_recvbuf,
                              const int* nbvalid27);

/* ---- several ranks (one GPU each): chunks partitioned along the space-filling-curve order ----
 * The host keeps ChunkMap / Balancer and hands over the rank boundaries (ChunkMap::get_rank,
 * chunkmap.cpp:156-164; Balancer::assign_initial, balancer.cpp:101-124).  Neighbours on other ranks
 * are then served by ONE message per (peer rank, mode) over NCCL send/recv instead of the
 * reference's 26 MPI messages per chunk and mode (Chunk::begin_bc_exchange / end_bc_exchange,
 * chunk.hpp:507-586; probe_bc_exchange, chunk.cpp:310-395); nixb200_domain_exchange_* /
 * _migrate_sort / _step include the cross-rank part once set_ranks + comm_init have been called.
 *
 * The plan (which slab goes to which rank, in the order both sides derive independently: sorted by
 * sender chunk id, then direction) is pure host logic and needs no device. */
typedef struct nixb200_plan nixb200_plan;
int nixb200_plan_create(const int* cdims /*[3]*/, const int* dims /*[3]*/, int nb, const int* coord,
                        int nrank, const int* boundary /*[nrank+1]*/, int rank, nixb200_plan** out);
int nixb200_plan_destroy(nixb200_plan* p);
int nixb200_plan_npeer(const nixb200_plan* p);
int nixb200_plan_peer(const nixb200_plan* p, int q, int* peer_rank, int* nsend, int* nrecv);
/* send3 [nsend][3] = (my chunk id, direction, cells); recv3 [nrecv][3] = (my chunk id, receive
 * slot, cells); entry j of my send list to rank r pairs with entry j of r's receive list from me */
int nixb200_plan_entries(const nixb200_plan* p, int q, int* send3, int* recv3);

/* boundary[rank], boundary[rank+1] must equal the domain's id_begin, id_end */
int nixb200_domain_set_ranks(nixb200_domain* d, int nrank, const int* boundary, int rank);
/* NCCL bootstrap: rank 0 calls comm_unique_id, the host broadcasts the 128 bytes (MPI_Bcast in a nix
 * application), every rank calls comm_init -- or hands over a communicator it already owns */
int nixb200_comm_unique_id(void* id128);
int nixb200_domain_comm_init(nixb200_domain* d, const void* id128);
int nixb200_domain_set_comm(nixb200_domain* d, void* nccl_comm);
/* a communicator the APPLICATION owns (one per run; handed to every domain it rebuilds after a rebalance with
 * nixb200_domain_set_comm, which never destroys it) */
int nixb200_comm_create(int nrank, int rank, const void* id128, int device, void** nccl_comm);
int nixb200_comm_destroy(void* nccl_comm);
int nixb200_device_count(int* n);
/* per step: ghost cells sent to other ranks per halo exchange; particles sent / received by the
 * last migrate */
int nixb200_domain_peer_traffic(nixb200_domain* d, int64_t* halo_cells_sent, int64_t* particles_sent,
                                int64_t* particles_received);

/* ---- chunks in the reference's wire format, made on the device; device-to-device rebalance ----
 * wire_size / wire_pack: the bytes that follow nix::Chunk::pack's own header (chunk.cpp:18-60) in the record of a
 *   PIC chunk: int order, int ns, uf[Mz][My][Mx][6], uj[Mz][My][Mx][4], then per species XtensorParticle::pack
 *   (xtensor_particle.hpp:128-169: 29 scalars, xu[Np_total][7], xv, gindex, pindex[Ng+1], pcount[Ng+1][8]; Np_total
 *   = ((Np + 128) / 128) * 128; xv and gindex -- scratch in the reference as well -- are zero).  `buffer` may be
 *   host or device memory.  Header + payload is a record the reference's own unpack reads (checkpoints,
 *   Balancer::sendrecv_chunk).
 * domain_rebalance: COLLECTIVE over the ranks of the communicator.  boundary = the new rank boundaries (the
 *   host's unchanged Balancer::assign, balancer.cpp:126-132).  Chunks that change owner travel GPU to GPU over
 *   NCCL in the wire format above, to rank-1 / rank+1 only (as Balancer::sendrecv_chunk, balancer.hpp:122-332:
 *   the old and the new range of a rank must overlap); chunks that stay are moved device to device into the
 *   re-sized arrays.  On return the handle covers [boundary[rank], boundary[rank+1]), rank tables and communicator
 *   are in place, counts are rebuilt.  Needs old + new arrays at once. */
int nixb200_chunk_wire_size(nixb200_domain* d, int k, int64_t* bytes);
/* the payload size from the chunk shape (Nz, Ny, Nx), the margin and the particle count of every species alone
 * (host logic, no device) */
int nixb200_wire_size_dims(const int* dims3, int nb, int ns, const int* np, int64_t* bytes);
/* the 175 scalar bytes XtensorParticle::pack opens a species with (xtensor_particle.hpp:130-159; host logic): offset3 /
 * gdims3 in cells as Chunk::set_global_context takes them (chunk.cpp:239-247), triples in (z, y, x) order */
int nixb200_wire_particle_header(const int* dims3, int nb, const double* del3, const int* offset3, const int* gdims3,
                                 double q, double m, int np, void* out175);
int nixb200_chunk_wire_pack(nixb200_domain* d, int k, void* buffer, int64_t bytes);
int nixb200_domain_rebalance(nixb200_domain* d, int nrank, const int* boundary, int rank);
/* host logic of it (no device needed): a rank that owned [b0, e0) and will own [b1, e1) sends [out0, out1) to rank-1 and
 * [out2, out3) to rank+1, receives [out4, out5) from rank-1 and [out6, out7) from rank+1, keeps [out8, out9) */
int nixb200_rebalance_moves(int b0, int e0, int b1, int e1, int* out10);

/* device-time accounting per phase (feeds Chunk::load; also bench.py's roofline).  Phases:
 * 0 push_deposit, 1 exchange_current, 2 exchange_field, 3 migrate+sort, 4 sort (count+sort only),
 * 5 the k_push launches of phase 0 alone, 6 its k_deposit launches alone (one call = one launch = one species),
 * 7 field solver (push_bfd + push_efd).
 * With profiling on, every phase call is bracketed by CUDA events on the domain's stream;
 * get_phase_ms synchronises, then returns and resets the accumulated milliseconds and call count. */
#define NIXB200_NPHASE 8
int nixb200_domain_set_profiling(nixb200_domain* d, int on);
int nixb200_domain_get_phase_ms(nixb200_domain* d, int phase, double* ms_sum, int* calls);

/* Chunk::get_total_load (chunk.hpp:189-192): device milliseconds of the last push_deposit */
int nixb200_domain_get_load(nixb200_domain* d, double* ms);
int64_t nixb200_domain_total_particles(nixb200_domain* d);

#ifdef __cplusplus
}
#endif
#endif
