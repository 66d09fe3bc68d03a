// demo_main.cpp -- the reference's own host classes (ChunkMap with its Gilbert space-filling curve,
// Chunk) driving the B200 path through host/nixb200_host.hpp, without MPI (single rank):
//
//   demo coord Cz Cy Cx                     print the id -> (cz,cy,cx) table of nix::ChunkMap
//   demo run <dir> Cz Cy Cx N order nb ns ppc_max steps
//        reads  <dir>/uf_<id>.bin, <dir>/xu_<id>_<is>.bin  (raw float64, the reference's layouts)
//        runs   `steps` x Application::push() worth of work in strict mode
//        writes <dir>/out_uj_<id>.bin, <dir>/out_xu_<id>_<is>.bin via Chunk staging (sync_host)
//        and    <dir>/pack_<id>.bin = GpuChunk::pack() bytes, unpacked again and compared in-process
//
// tests/test_host_cpp.py compares the outputs with the CPU oracle (bit-exact particles).
#include "nixb200_host.hpp"
#include "balancer.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

using namespace nixb200host;

static std::vector<double> read_bin(const std::string& path)
{
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return {};
  size_t              n = (size_t)f.tellg() / sizeof(double);
  std::vector<double> v(n);
  f.seekg(0);
  f.read(reinterpret_cast<char*>(v.data()), n * sizeof(double));
  return v;
}

static void write_bin(const std::string& path, const double* p, size_t n)
{
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(p), n * sizeof(double));
}

int main(int argc, char** argv)
{
  if (argc >= 5 && std::string(argv[1]) == "coord") {
    nix::ChunkMap cm(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
    int           n = atoi(argv[2]) * atoi(argv[3]) * atoi(argv[4]);
    for (int id = 0; id < n; id++) {
      auto [cz, cy, cx] = cm.get_coordinate(id);
      std::printf("%d %d %d\n", cz, cy, cx);
    }
    return 0;
  }
  if (argc >= 4 && std::string(argv[1]) == "assign") {
    // demo assign <nrank> <initial|step> load0 load1 ... [-- b0 b1 ... b_nrank]: the reference's own Balancer
    // (balancer.cpp:8-132) on a load vector -> the new rank boundaries (golden vectors for nix_b200/balancer.py)
    const int            nrank = atoi(argv[2]);
    const bool           init  = std::string(argv[3]) == "initial";
    std::vector<double>  load;
    std::vector<int>     boundary;
    bool                 second = false;
    for (int i = 4; i < argc; i++) {
      if (std::string(argv[i]) == "--") second = true;
      else if (second) boundary.push_back(atoi(argv[i]));
      else load.push_back(atof(argv[i]));
    }
    nix::Balancer bal((int)load.size());
    for (size_t i = 0; i < load.size(); i++) bal.load(i) = load[i];
    std::vector<int> out = init ? bal.assign_initial(nrank) : bal.assign(boundary);
    for (int b : out) std::printf("%d ", b);
    std::printf("\n");
    return 0;
  }
  if (argc >= 8 && std::string(argv[1]) == "wiresize") {
    // demo wiresize Nz Ny Nx nb order np_0 [np_1 ...] : bytes a PIC chunk appends to nix::Chunk::pack's own header,
    // measured by the reference's OWN pack() in query mode (chunk.cpp:18-60, xtensor_particle.hpp:128-169) on
    // containers sized for exactly np particles.  Host only: no device, no domain.
    const int dz = atoi(argv[2]), dy = atoi(argv[3]), dx = atoi(argv[4]), nb = atoi(argv[5]), order = atoi(argv[6]);
    GpuInterface factory;
    auto         c = factory.create_chunk(nix::Dims3D{dz, dy, dx}, nix::Bool3D{true, true, true}, 0);
    auto*        g = static_cast<GpuChunk*>(c.get());
    int          offset[3] = {0, 0, 0}, gdims[3] = {dz, dy, dx};
    g->set_boundary_margin(nb);
    g->set_global_context(offset, gdims);
    g->set_coordinate(1.0, 1.0, 1.0);
    g->allocate_staging(order, nb, {}, 0);
    for (int i = 7; i < argc; i++) {
      const int np = atoi(argv[i]);
      auto      p  = std::make_shared<nix::XtensorParticle>(np, *g);
      p->Np        = np;
      g->up.push_back(p);
    }
    const int all = g->pack(nullptr, 0), hdr = g->nix::Chunk::pack(nullptr, 0);
    std::printf("%d %d\n", hdr, all - hdr);
    return 0;
  }
  if (argc == 17 && std::string(argv[1]) == "wirehdr") {
    // demo wirehdr Nz Ny Nx nb delz dely delx oz oy ox gz gy gx q np : the first 175 bytes of the reference's OWN
    // XtensorParticle::pack (xtensor_particle.hpp:130-159) for a chunk of that geometry, as hex.  Host only.
    const int    d[3] = {atoi(argv[2]), atoi(argv[3]), atoi(argv[4])}, nb = atoi(argv[5]);
    const double del[3] = {atof(argv[6]), atof(argv[7]), atof(argv[8])};
    int          offset[3] = {atoi(argv[9]), atoi(argv[10]), atoi(argv[11])}, gdims[3] = {atoi(argv[12]), atoi(argv[13]), atoi(argv[14])};
    const double q = atof(argv[15]);
    const int    np = atoi(argv[16]);
    GpuInterface factory;
    auto         c = factory.create_chunk(nix::Dims3D{d[0], d[1], d[2]}, nix::Bool3D{true, true, true}, 0);
    auto*        g = static_cast<GpuChunk*>(c.get());
    g->set_boundary_margin(nb);
    g->set_global_context(offset, gdims);
    g->set_coordinate(del[0], del[1], del[2]);
    nix::XtensorParticle p(np, *g);
    p.Np = np;
    p.q  = q;
    p.m  = 25.0;
    std::vector<uint8_t> buf(p.pack(nullptr, 0));
    p.pack(buf.data(), 0);
    for (int i = 0; i < 175; i++) std::printf("%02x", buf[i]);
    std::printf("\n");
    return 0;
  }
  if (argc < 12 || std::string(argv[1]) != "run") {
    std::fprintf(stderr, "usage: demo coord Cz Cy Cx | demo run dir Cz Cy Cx N order nb ns ppc_max steps\n");
    return 2;
  }
  const std::string dir = argv[2];
  const int cdims[3] = {atoi(argv[3]), atoi(argv[4]), atoi(argv[5])};
  const int N = atoi(argv[6]), order = atoi(argv[7]), nb = atoi(argv[8]), ns = atoi(argv[9]);
  const int npmax = atoi(argv[10]), steps = atoi(argv[11]);
  const int nchunk = cdims[0] * cdims[1] * cdims[2];
  try {
    nix::ChunkMap    chunkmap(cdims[0], cdims[1], cdims[2]);
    std::vector<int> boundary = {0, nchunk};
    chunkmap.set_rank_boundary(boundary);
    std::vector<SpeciesSpec> species;
    for (int is = 0; is < ns; is++) species.push_back(is == 0 ? SpeciesSpec{-1.0, 1.0} : SpeciesSpec{1.0, 25.0});

    GpuInterface                             factory;
    std::vector<std::unique_ptr<nix::Chunk>> chunks;
    int gdims[3] = {cdims[0] * N, cdims[1] * N, cdims[2] * N};
    for (int id = 0; id < nchunk; id++) {
      auto c  = factory.create_chunk(nix::Dims3D{N, N, N}, nix::Bool3D{true, true, true}, id);
      auto* g = static_cast<GpuChunk*>(c.get());
      auto [cz, cy, cx] = chunkmap.get_coordinate(id);
      int offset[3]     = {cz * N, cy * N, cx * N};
      g->set_boundary_margin(nb);
      g->set_global_context(offset, gdims);
      g->set_coordinate(1.0, 1.0, 1.0);
      g->allocate_staging(order, nb, species, npmax);
      auto uf = read_bin(dir + "/uf_" + std::to_string(id) + ".bin");
      if (uf.size() != g->uf.size()) throw std::runtime_error("bad uf file for chunk " + std::to_string(id));
      std::copy(uf.begin(), uf.end(), g->uf.data());
      for (int is = 0; is < ns; is++) {
        auto xu = read_bin(dir + "/xu_" + std::to_string(id) + "_" + std::to_string(is) + ".bin");
        int  np = (int)(xu.size() / 7);
        if (np > g->up[is]->Np_total) g->up[is]->resize(np);
        std::copy(xu.begin(), xu.end(), g->up[is]->xu.data());
        g->up[is]->Np = np;
      }
      chunks.push_back(std::move(c));
    }
    auto dom = make_domain(chunks, chunkmap, cdims, nb, order, species, 1.0, 0, /*strict_fp=*/true, 2.0);
    check(nixb200_domain_exchange_field(dom->h), "exchange_field");
    check(nixb200_domain_sort(dom->h), "sort");
    for (int s = 0; s < steps; s++) check(nixb200_domain_step(dom->h, 0.5), "step");
    int err = 0;
    check(nixb200_domain_check(dom->h, &err), "check");
    if (err) throw std::runtime_error("device error bits " + std::to_string(err));
    long long total = 0;
    for (int id = 0; id < nchunk; id++) {
      auto* g          = static_cast<GpuChunk*>(chunks[id].get());
      g->host_is_newer = false;
      g->host_synced   = false;
      g->sync_host();
      write_bin(dir + "/out_uj_" + std::to_string(id) + ".bin", g->uj.data(), g->uj.size());
      for (int is = 0; is < ns; is++) {
        write_bin(dir + "/out_xu_" + std::to_string(id) + "_" + std::to_string(is) + ".bin", g->up[is]->xu.data(),
                  (size_t)g->up[is]->Np * 7);
        total += g->up[is]->Np;
      }
      // Chunk::pack in query mode, then for real, then unpack into a fresh chunk (what
      // Balancer::sendrecv_chunk and StateHandler do) and compare
      int                  bytes = g->pack(nullptr, 0);
      std::vector<uint8_t> buf(bytes);
      if (g->pack(buf.data(), 0) != bytes) throw std::runtime_error("pack size mismatch");
      auto  c2 = factory.create_chunk(nix::Dims3D{N, N, N}, nix::Bool3D{true, true, true}, 0);
      auto* g2 = static_cast<GpuChunk*>(c2.get());
      if (g2->unpack(buf.data(), 0) != bytes) throw std::runtime_error("unpack size mismatch");
      bool same = g2->get_id() == id && g2->order == order && (int)g2->up.size() == ns && g2->uf == g->uf && g2->uj == g->uj;
      for (int is = 0; same && is < ns; is++) same = g2->up[is]->Np == g->up[is]->Np && g2->up[is]->xu == g->up[is]->xu;
      if (!same) throw std::runtime_error("pack/unpack round trip differs for chunk " + std::to_string(id));
      // the DEVICE-made payload (nixb200_chunk_wire_pack) behind the reference's own header, read back by the
      // reference's own unpack (nix::Chunk::unpack + XtensorParticle::unpack): same chunk
      int     hdr = g->nix::Chunk::pack(nullptr, 0);
      int64_t pay = 0;
      check(nixb200_chunk_wire_size(dom->h, id, &pay), "wire_size");
      std::vector<uint8_t> rec((size_t)hdr + pay);
      g->nix::Chunk::pack(rec.data(), 0);
      check(nixb200_chunk_wire_pack(dom->h, id, rec.data() + hdr, pay), "wire_pack");
      auto  c3 = factory.create_chunk(nix::Dims3D{N, N, N}, nix::Bool3D{true, true, true}, 0);
      auto* g3 = static_cast<GpuChunk*>(c3.get());
      if (g3->unpack(rec.data(), 0) != hdr + pay) throw std::runtime_error("device wire record: size mismatch");
      bool wsame = g3->get_id() == id && g3->order == order && (int)g3->up.size() == ns && g3->uf == g->uf && g3->uj == g->uj;
      for (int is = 0; wsame && is < ns; is++) {
        auto& a = *g3->up[is];
        auto& b = *g->up[is];
        wsame   = a.Np == b.Np && a.Ng == b.Ng && a.q == b.q && a.m == b.m && a.xmin == b.xmin && a.zmax == b.zmax &&
                a.delx == b.delx && a.Lbx == b.Lbx && a.Ubz == b.Ubz && a.xmax_global == b.xmax_global;
        for (int ip = 0; wsame && ip < a.Np; ip++)
          for (int c = 0; c < 7; c++) wsame = wsame && std::memcmp(&a.xu(ip, c), &b.xu(ip, c), 8) == 0;
        for (int ii = 0; wsame && ii <= a.Ng; ii++) wsame = a.pindex(ii) == b.pindex(ii);
      }
      if (!wsame) throw std::runtime_error("device wire record differs for chunk " + std::to_string(id));
    }
    std::printf("ok chunks=%d particles=%lld launches=%lld\n", nchunk, total, (long long)nixb200_launch_count());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "demo failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
