// nixb200_host.hpp -- the C++ host side of the drop-in: what a nix application instantiates so that
// its per-chunk PIC step runs on B200 through libnixb200.so (include/nixb200.h).
//
// It keeps the reference's own interfaces (north star: "the C++ host keeps the Chunk / ChunkMap /
// Application / Balancer interfaces and the SFC chunk ordering"):
//
//   GpuChunk       : nix::Chunk                  chunk.hpp:100-432   (factory product, pack/unpack, load)
//   GpuInterface   : nix::Application::Interface application.hpp:23-66 (create_chunk -> GpuChunk)
//   GpuApplication : nix::Application            application.hpp:69-396 (push() override, rebalance hook)
//
// and calls ONLY the C ABI.  The reference drives every chunk separately (26 MPI messages per chunk
// and mode); here all chunks of the rank live in ONE device-resident nixb200_domain and
// Application::push() is a single nixb200_domain_step().  A GpuChunk is therefore a thin handle
// (domain, local index) plus host staging in the reference's own containers (xtensor uf/uj,
// XtensorParticle), which is what Chunk::pack / unpack (checkpoints, Balancer::sendrecv_chunk) and the
// diagnostics read and write -- byte-compatible with the reference's formats because the
// reference's own pack() writes them.
//
// Compile with the reference on the include path:  -I<nix> -I<nix>/thirdparty -I<repo>/include
#pragma once

#include "application.hpp"
#include "chunk.hpp"
#include "chunkmap.hpp"
#include "diag.hpp"
#include "xtensor_particle.hpp"

#include "nixb200.h"

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace nixb200host
{
using nix::float64;
using nix::json;

/// ERROR-and-carry-on is the reference's convention on the hot path; set-up failures
/// are fatal there (MPI_Abort), here they throw.
inline void check(int rc, const char* what)
{
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + nixb200_last_error());
}

struct SpeciesSpec {
  float64 q, m;
};

/// RAII owner of one nixb200_domain = the chunks [id_begin, id_end) of this rank on one GPU
class GpuDomain
{
public:
  nixb200_domain* h = nullptr;
  int             id_begin = 0, id_end = 0;

  GpuDomain(nix::ChunkMap& chunkmap, const int cdims[3], const int dims[3], int nb, int order,
            const std::vector<SpeciesSpec>& species, float64 delz, float64 dely, float64 delx, float64 cc,
            int id_begin_, int id_end_, int device, bool strict_fp = false, double capacity_factor = 1.25,
            int pusher = NIXB200_PUSH_BORIS)
      : id_begin(id_begin_), id_end(id_end_)
  {
    const int        ncid = cdims[0] * cdims[1] * cdims[2];
    std::vector<int> coord(3 * ncid);
    for (int id = 0; id < ncid; id++) { // ChunkMap::get_coordinate = the Gilbert curve of sfc.cpp
      auto [cz, cy, cx] = chunkmap.get_coordinate(id);
      coord[3 * id + 0] = cz;
      coord[3 * id + 1] = cy;
      coord[3 * id + 2] = cx;
    }
    nixb200_domain_desc desc{};
    for (int a = 0; a < 3; a++) {
      desc.cdims[a] = cdims[a];
      desc.dims[a]  = dims[a];
    }
    desc.nb              = nb;
    desc.order           = order;
    desc.ns              = (int)species.size();
    desc.del[0]          = delz;
    desc.del[1]          = dely;
    desc.del[2]          = delx;
    desc.cc              = cc;
    desc.id_begin        = id_begin;
    desc.id_end          = id_end;
    desc.device          = device;
    desc.strict_fp       = strict_fp ? 1 : 0;
    desc.capacity_factor = capacity_factor;
    desc.pusher          = pusher; // push_boris / push_vay / push_higuera_cary (primitives.hpp:165-253)
    desc.fp32            = 0;      // the host mirror keeps the reference's real type
    std::vector<double> q, m;
    for (auto& s : species) {
      q.push_back(s.q);
      m.push_back(s.m);
    }
    check(nixb200_domain_create(&desc, coord.data(), q.data(), m.data(), &h), "nixb200_domain_create");
  }
  GpuDomain(const GpuDomain&)            = delete;
  GpuDomain& operator=(const GpuDomain&) = delete;
  ~GpuDomain()
  {
    if (h) nixb200_domain_destroy(h);
  }

  /// several ranks: boundary = ChunkMap::get_rank_boundary(); id128 = the NCCL unique id every rank
  /// received from rank 0 (MPI_Bcast of nixb200_comm_unique_id's 128 bytes)
  void set_ranks(std::vector<int> boundary, int rank, const void* id128)
  {
    check(nixb200_domain_set_ranks(h, (int)boundary.size() - 1, boundary.data(), rank), "nixb200_domain_set_ranks");
    if ((int)boundary.size() - 1 > 1) check(nixb200_domain_comm_init(h, id128), "nixb200_domain_comm_init");
  }
};

/// Chunk whose fields and particles live in a GpuDomain.  Host staging uses the reference's own
/// containers so that pack()/unpack() and diagnostics keep the reference's byte formats.
class GpuChunk : public nix::Chunk
{
public:
  using ParticlePtr = std::shared_ptr<nix::XtensorParticle>;

  int                          order = 2;
  xt::xtensor<float64, 4>      uf; ///< [Mz][My][Mx][6] host staging of E/B
  xt::xtensor<float64, 4>      uj; ///< [Mz][My][Mx][4] host staging of rho/J
  std::vector<ParticlePtr>     up; ///< host staging of the species (reference container)
  std::shared_ptr<GpuDomain>   domain;
  int                          local = -1; ///< index inside the domain
  bool                         host_is_newer = true;  ///< staging holds data the device has not seen
  bool                         host_synced   = false; ///< staging equals the device state (no step since the last download)

  GpuChunk(nix::Dims3D dims, nix::Bool3D has_dim, int id = 0) : nix::Chunk(dims, has_dim, id)
  {
  }

  int get_order() const // duck-typed by XtensorHaloParticle3D (xtensor_halo3d.hpp:545)
  {
    return order;
  }

  /// allocate the staging; the downstream application fills uf / up[is]->xu in its setup()
  void allocate_staging(int order_, int nb, const std::vector<SpeciesSpec>& species, int np_required)
  {
    order = order_;
    set_boundary_margin(nb);
    size_t mz = dims[0] + 2 * nb, my = dims[1] + 2 * nb, mx = dims[2] + 2 * nb;
    uf.resize({mz, my, mx, 6ul});
    uj.resize({mz, my, mx, 4ul});
    uf.fill(0);
    uj.fill(0);
    up.clear();
    for (auto& s : species) {
      auto p = std::make_shared<nix::XtensorParticle>(np_required, *this);
      p->q   = s.q;
      p->m   = s.m;
      p->Np  = 0;
      up.push_back(p);
    }
    load.assign(1, 0.0);
  }

  int64_t get_size_byte() const override
  {
    int64_t size = uf.size() * sizeof(float64) + uj.size() * sizeof(float64);
    for (auto& p : up) size += p->get_size_byte();
    return size;
  }

  /// device -> host staging (before pack, diagnostics, rebalance)
  void sync_host()
  {
    if (!domain || local < 0 || host_is_newer || host_synced) return;
    check(nixb200_chunk_field_download(domain->h, local, NIXB200_FIELD_UF, uf.data()), "field_download");
    check(nixb200_chunk_field_download(domain->h, local, NIXB200_FIELD_UJ, uj.data()), "field_download");
    for (size_t is = 0; is < up.size(); is++) {
      int64_t np = 0;
      check(nixb200_chunk_get_particles(domain->h, local, (int)is, nullptr, 0, &np), "get_particles");
      if (np > up[is]->Np_total - 1) up[is]->resize((int)np);
      if (np > 0) check(nixb200_chunk_get_particles(domain->h, local, (int)is, up[is]->xu.data(), np, &np), "get_particles");
      up[is]->Np = (int)np;
      check(nixb200_chunk_get_pindex(domain->h, local, (int)is, up[is]->pindex.data()), "get_pindex");
    }
    host_synced = true;
  }

  // Chunk::pack / unpack (chunk.cpp:18-116): base header first, then this chunk's arrays in the
  // reference's formats (XtensorParticle::pack, xtensor_particle.hpp:128-218)
  int pack(void* buffer, int address) override
  {
    using nix::memcpy_count;
    sync_host();
    int ns = (int)up.size();
    address = nix::Chunk::pack(buffer, address);
    address += memcpy_count(buffer, &order, sizeof(int), address, 0);
    address += memcpy_count(buffer, &ns, sizeof(int), address, 0);
    address += memcpy_count(buffer, uf.data(), uf.size() * sizeof(float64), address, 0);
    address += memcpy_count(buffer, uj.data(), uj.size() * sizeof(float64), address, 0);
    for (auto& p : up) address = p->pack(buffer, address);
    return address;
  }

  int unpack(void* buffer, int address) override
  {
    using nix::memcpy_count;
    int ns = 0;
    address = nix::Chunk::unpack(buffer, address);
    address += memcpy_count(&order, buffer, sizeof(int), 0, address);
    address += memcpy_count(&ns, buffer, sizeof(int), 0, address);
    size_t mz = dims[0] + 2 * boundary_margin, my = dims[1] + 2 * boundary_margin, mx = dims[2] + 2 * boundary_margin;
    uf.resize({mz, my, mx, 6ul});
    uj.resize({mz, my, mx, 4ul});
    address += memcpy_count(uf.data(), buffer, uf.size() * sizeof(float64), 0, address);
    address += memcpy_count(uj.data(), buffer, uj.size() * sizeof(float64), 0, address);
    up.resize(ns);
    for (int is = 0; is < ns; is++) {
      up[is]  = std::make_shared<nix::XtensorParticle>();
      address = up[is]->unpack(buffer, address);
    }
    host_is_newer = true; // the next push() uploads it into the rank's domain
    domain.reset();
    local = -1;
    return address;
  }

  void set_load_value(float64 v)
  {
    load.assign(1, v);
  }

  // the exchange of all chunk pairs happens inside nixb200_domain_step: the per-chunk hooks of the
  // reference's main loop are satisfied trivially
  void set_boundary_pack(int) override {}
  void set_boundary_unpack(int) override {}
  void set_boundary_begin(int) override {}
  void set_boundary_end(int) override {}
  bool set_boundary_probe(int, bool) override { return true; }
};

/// XtensorPacker3D (xtensor_packer3d.hpp:15-140) for chunks whose data live on the device: what a downstream
/// Packer functor of nix::ChunkWriter (diag/chunk_writer.hpp:118-146) calls instead of sync_host() + the host
/// packer.  Same conventions -- the return value is `address + bytes`, buffer == nullptr only queries -- and the
/// same bytes: the device packers (csrc/diag.cu) reproduce the reference's colocation, block averages and
/// rounding bit for bit, so only the PACKED output crosses PCIe.  The chunk must be resident in its rank's
/// domain (after the first push(), or after an explicit GpuApplication::ensure_domain()).
class GpuPacker3D
{
  static nixb200_domain* handle(GpuChunk& c)
  {
    if (!c.domain || c.local < 0 || c.host_is_newer)
      throw std::runtime_error("GpuPacker3D: chunk is not resident on the device (call ensure_domain() first)");
    return c.domain->h;
  }

public:
  /// pack_field (xtensor_packer3d.hpp:62-82): E/B colocated at the cell centres, block-averaged by `decimate`
  size_t pack_field(GpuChunk& c, int decimate, uint8_t* buffer, int address)
  {
    int64_t n = 0;
    check(nixb200_chunk_pack_field(handle(c), c.local, decimate, nullptr, &n), "pack_field");
    if (buffer) check(nixb200_chunk_pack_field(handle(c), c.local, decimate, reinterpret_cast<double*>(buffer + address), &n), "pack_field");
    return sizeof(float64) * (size_t)n + address;
  }

  /// pack_moment (:84-104): block averages of uj (which = 0) or of the ns x 14 moments (which = 1; after
  /// nixb200_domain_deposit_moment)
  size_t pack_moment(GpuChunk& c, int which, int decimate, uint8_t* buffer, int address)
  {
    int64_t n = 0;
    check(nixb200_chunk_pack_moment(handle(c), c.local, which, decimate, nullptr, &n), "pack_moment");
    if (buffer) check(nixb200_chunk_pack_moment(handle(c), c.local, which, decimate, reinterpret_cast<double*>(buffer + address), &n), "pack_moment");
    return sizeof(float64) * (size_t)n + address;
  }

  /// pack_tracer (:122-140): the particles of species `is` with a negative 64-bit id, in container order
  size_t pack_tracer(GpuChunk& c, int is, uint8_t* buffer, int address)
  {
    int64_t np = 0;
    check(nixb200_chunk_pack_tracer(handle(c), c.local, is, nullptr, 0, &np), "pack_tracer");
    if (buffer && np > 0) check(nixb200_chunk_pack_tracer(handle(c), c.local, is, reinterpret_cast<double*>(buffer + address), np, &np), "pack_tracer");
    return sizeof(float64) * 7 * (size_t)np + address;
  }
};

/// Application::Interface whose factory makes GpuChunks (application.hpp:45-48); used at set-up, on
/// rebalance receive (balancer.hpp:216) and on checkpoint load (statehandler.hpp:297)
class GpuInterface : public nix::Application::Interface
{
public:
  PtrChunk create_chunk(nix::Dims3D dims, nix::Bool3D has_dim, int id) override
  {
    return std::make_unique<GpuChunk>(dims, has_dim, id);
  }
};

/// Build (or rebuild) the rank's domain from its chunk vector and upload every chunk's staging.
/// `chunks` are the rank-local chunks in ascending id order (ChunkVector keeps them sorted).  Chunks that
/// still live in a previous domain are brought down to their staging first (lazily: only here and in
/// GpuChunk::pack, never per step).
template <typename ChunkVec>
std::shared_ptr<GpuDomain> make_domain(ChunkVec& chunks, nix::ChunkMap& chunkmap, const int cdims[3], int nb, int order,
                                       const std::vector<SpeciesSpec>& species, float64 cc, int device, bool strict_fp,
                                       double capacity_factor = 1.25, int pusher = NIXB200_PUSH_BORIS)
{
  if (chunks.size() == 0) return nullptr;
  for (auto& c : chunks) static_cast<GpuChunk*>(c.get())->sync_host();
  auto* first = static_cast<GpuChunk*>(chunks.front().get());
  auto* last  = static_cast<GpuChunk*>(chunks.back().get());
  std::vector<int> nd = first->get_dims();
  int dims[3]         = {nd[0], nd[1], nd[2]};
  auto dom = std::make_shared<GpuDomain>(chunkmap, cdims, dims, nb, order, species, first->get_delz(), first->get_dely(),
                                         first->get_delx(), cc, first->get_id(), last->get_id() + 1, device, strict_fp,
                                         capacity_factor, pusher);
  const int ns = (int)species.size();
  for (int k = 0; k < (int)chunks.size(); k++) {
    auto* c = static_cast<GpuChunk*>(chunks[k].get());
    if (c->get_id() != dom->id_begin + k) throw std::runtime_error("chunk ids of a rank must be contiguous");
    c->domain = dom;
    c->local  = k;
    check(nixb200_chunk_field_upload(dom->h, k, NIXB200_FIELD_UF, c->uf.data()), "field_upload");
  }
  for (int is = 0; is < ns; is++) {
    std::vector<int64_t> np(chunks.size());
    int64_t              total = 0;
    for (size_t k = 0; k < chunks.size(); k++) total += (np[k] = static_cast<GpuChunk*>(chunks[k].get())->up[is]->Np);
    std::vector<double> flat((size_t)total * 7);
    size_t              pos = 0;
    for (size_t k = 0; k < chunks.size(); k++) {
      auto* c = static_cast<GpuChunk*>(chunks[k].get());
      std::copy(c->up[is]->xu.data(), c->up[is]->xu.data() + np[k] * 7, flat.begin() + pos);
      pos += np[k] * 7;
    }
    check(nixb200_domain_set_particles(dom->h, is, flat.data(), np.data()), "set_particles");
  }
  for (auto& c : chunks) {
    static_cast<GpuChunk*>(c.get())->host_is_newer = false;
    static_cast<GpuChunk*>(c.get())->host_synced   = false; // (sort may drop / reorder particles)
  }
  return dom;
}

/// nix::Application with the PIC step on the GPU.  A downstream application derives from this
/// instead of nix::Application, fills the staging of its GpuChunks in Chunk::setup(), and keeps
/// everything else (config, diagnostics, checkpoints, balancer) unchanged.
class GpuApplication : public nix::Application
{
protected:
  std::shared_ptr<GpuDomain> domain;
  std::vector<SpeciesSpec>   species;
  int                        order = 2, nb = 2, device = -1; ///< device < 0: node-local rank modulo the device count
  float64                    cc    = 1.0;
  bool                       strict_fp    = true;  ///< bit-exact with the reference's scalar templates (1.8 % slower)
  bool                       field_solver = false; ///< true: nixb200_domain_step_em (E/B never leave the device)
  float64                    cfj          = 1.0;
  int                        pusher       = NIXB200_PUSH_BORIS;
  void*                      nccl_comm    = nullptr; ///< created once, handed to every rebuilt domain
  float64                    cell_weight  = 12.0; ///< load of one cell and species in particle-equivalents
  long                       npush = 0, nrebuild = 0;

public:
  GpuApplication(int argc, char** argv, PtrInterface interface = std::make_shared<GpuInterface>())
      : nix::Application(argc, argv, interface)
  {
  }

  ~GpuApplication() override
  {
    domain.reset();
    if (nccl_comm) nixb200_comm_destroy(nccl_comm);
  }

  /// node-local rank -> device (one rank per GPU)
  virtual int select_device()
  {
    if (device >= 0) return device;
    int ndev = 0;
    check(nixb200_device_count(&ndev), "nixb200_device_count");
    MPI_Comm node;
    int      local = 0;
    MPI_Comm_split_type(MPI_COMM_WORLD, MPI_COMM_TYPE_SHARED, thisrank, MPI_INFO_NULL, &node);
    MPI_Comm_rank(node, &local);
    MPI_Comm_free(&node);
    return device = local % ndev;
  }

  /// (Re)build the device-resident domain when the chunk set of ANY rank changed.  Every decision that
  /// leads to a collective call is itself collective: the staleness flag is all-reduced and the NCCL
  /// bootstrap runs on every rank, whether or not it owns chunks.
  virtual void ensure_domain()
  {
    int stale = !domain && chunkvec.size() > 0;
    for (auto& c : chunkvec) stale = stale || static_cast<GpuChunk*>(c.get())->domain != domain;
    MPI_Allreduce(MPI_IN_PLACE, &stale, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
    if (!stale) return;
    const int dev = select_device();
    if (nprocess > 1 && nccl_comm == nullptr) {
      unsigned char id[128] = {0};
      if (thisrank == 0) check(nixb200_comm_unique_id(id), "comm_unique_id");
      MPI_Bcast(id, 128, MPI_BYTE, 0, MPI_COMM_WORLD);
      check(nixb200_comm_create(nprocess, thisrank, id, dev, &nccl_comm), "nixb200_comm_create");
    }
    int cd[3] = {cdims[0], cdims[1], cdims[2]};
    domain    = make_domain(chunkvec, *chunkmap, cd, nb, order, species, cc, dev, strict_fp, 1.25, pusher);
    nrebuild++;
    if (domain) {
      auto boundary = chunkmap->get_rank_boundary();
      check(nixb200_domain_set_ranks(domain->h, nprocess, boundary.data(), thisrank), "nixb200_domain_set_ranks");
      if (nprocess > 1) check(nixb200_domain_set_comm(domain->h, nccl_comm), "nixb200_domain_set_comm");
      check(nixb200_domain_exchange_field(domain->h), "exchange_field");
      check(nixb200_domain_sort(domain->h), "sort");
    }
  }

  /// Application::push() (application.hpp:343-346): one step of every local chunk
  void push() override
  {
    ensure_domain();
    npush++;
    if (!domain) return;
    const float64 delt = cfgparser->get_delt();
    if (field_solver) check(nixb200_domain_step_em(domain->h, delt, cfj), "nixb200_domain_step_em");
    else check(nixb200_domain_step(domain->h, delt), "nixb200_domain_step");
    int err = 0;
    check(nixb200_domain_check(domain->h, &err), "nixb200_domain_check");
    if (err & NIXB200_ERR_CFL) ERROR << tfm::format("step[%d] a particle moved more than one cell (c*delt > delh)", curstep);
    if (err & NIXB200_ERR_UNSORTED) ERROR << tfm::format("step[%d] particle container was not cell-sorted", curstep);
    assert_mpi((err & NIXB200_ERR_CAPACITY) == 0, "particle storage overflowed on the device (raise capacity_factor)");
    // Chunk::load feeds the balancer (chunk.hpp:177-192): device time of the push, shared among the chunks in
    // proportion to (particles + cell_weight x cells x species): a bin costs the kernels about as much as 12
    // particles (per-bin reduction and flush of the deposit, the 8 keys per cell of the sort; measured,
    // profiles/r02p_cfg5_n2.json), which matters once there are fewer than ~30 particles per cell
    double ms = 0;
    nixb200_domain_get_load(domain->h, &ms);
    std::vector<int64_t> np(chunkvec.size());
    std::vector<double>  w(chunkvec.size(), 0.0);
    double               total = 0;
    for (size_t is = 0; is < species.size(); is++) {
      check(nixb200_domain_get_np(domain->h, (int)is, np.data()), "get_np");
      for (size_t k = 0; k < np.size(); k++) {
        auto         nd    = chunkvec[k]->get_dims();
        const double cells = cell_weight * nd[0] * nd[1] * nd[2];
        w[k] += (double)np[k] + cells;
        total += (double)np[k] + cells;
      }
    }
    for (size_t k = 0; k < chunkvec.size(); k++) {
      auto* c          = static_cast<GpuChunk*>(chunkvec[k].get());
      c->host_is_newer = false;
      c->host_synced   = false;
      c->set_load_value(total > 0 ? ms * w[k] / total : ms / chunkvec.size());
    }
  }

  /// Application::rebalance() (application.cpp:337-381) ships chunks through Chunk::pack / unpack
  /// (Balancer::sendrecv_chunk, balancer.hpp:161-222).  GpuChunk::pack brings ITS chunk down from the device
  /// on demand, so nothing is downloaded on the steps -- or for the chunks -- that do not move; the domain is
  /// rebuilt by the next push() only if the rank boundaries actually changed.
  bool rebalance() override
  {
    const auto before = chunkmap->get_rank_boundary();
    const bool ran    = nix::Application::rebalance();
    if (ran && chunkmap->get_rank_boundary() != before) {
      for (auto& c : chunkvec) static_cast<GpuChunk*>(c.get())->sync_host(); // those that stay: staged for the re-upload
      for (auto& c : chunkvec) {
        auto* g   = static_cast<GpuChunk*>(c.get());
        g->host_is_newer = true;
        g->domain.reset();
        g->local = -1;
      }
      domain.reset();
    }
    return ran;
  }

  /// everything back in the staging containers (before diagnostics that read Chunk data, before save)
  void sync_host_all()
  {
    for (auto& c : chunkvec) static_cast<GpuChunk*>(c.get())->sync_host();
  }
};
} // namespace nixb200host
