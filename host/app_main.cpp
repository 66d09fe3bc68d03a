// app_main.cpp -- a complete nix application on the B200 path: nix::Application::main() of the REFERENCE
// (application.cpp:46-87: initialize -> setup_chunks -> loop { diagnostic, push, rebalance, take_log,
// increment_time } -> finalize) drives GpuApplication::push() / rebalance() of host/nixb200_host.hpp.
// Built against the reference's own application.cpp / balancer.cpp / chunk.cpp / chunkmap.cpp / sfc.cpp /
// nixio.cpp where they lie, with the single-process MPI stand-in of host/stub when no MPI is installed.
//
//   app_main -c config.json --tmax T
//
// config.json: the reference's layout (unittest/test_application.cpp:18-57) plus
//   "application": { "option": { "input": "<dir>", "order": 2, "nb": 2, "strict": true, "field_solver": false,
//                                "cfj": 1.0, "species": [[q, m], ...], "np_max": N } }
// Every chunk reads <input>/uf_<id>.bin and <input>/xu_<id>_<is>.bin in its setup() (raw float64, the
// reference's array layouts) and the application writes <input>/out_xu_<id>_<is>.bin, out_uf_<id>.bin,
// out_uj_<id>.bin and app_summary.json from finalize().  tests/test_host_cpp.py compares them with the oracle.
#include "nixb200_host.hpp"

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

class FileChunk : public GpuChunk
{
public:
  using GpuChunk::GpuChunk;

  // Chunk::setup (chunk.hpp:146-150), called by Application::setup_chunks_init (application.cpp:316-320)
  // with config = "parameter" + "option"
  void setup(json& config) override
  {
    json        opt   = config["option"];
    float64     delh  = config.value("delh", 1.0);
    std::string input = opt.value("input", std::string("."));
    int         nbm   = opt.value("nb", 2);
    std::vector<SpeciesSpec> species;
    for (auto& s : opt["species"]) species.push_back(SpeciesSpec{s[0].get<double>(), s[1].get<double>()});
    set_boundary_margin(nbm);
    set_coordinate(delh, delh, delh);
    allocate_staging(opt.value("order", 2), nbm, species, opt.value("np_max", 1024));
    auto ufv = read_bin(input + "/uf_" + std::to_string(myid) + ".bin");
    if (ufv.size() != uf.size()) throw std::runtime_error("bad uf file for chunk " + std::to_string(myid));
    std::copy(ufv.begin(), ufv.end(), uf.data());
    for (size_t is = 0; is < species.size(); is++) {
      auto xu = read_bin(input + "/xu_" + std::to_string(myid) + "_" + std::to_string(is) + ".bin");
      int  np = (int)(xu.size() / 7);
      if (np > up[is]->Np_total) up[is]->resize(np);
      std::copy(xu.begin(), xu.end(), up[is]->xu.data());
      up[is]->Np = np;
    }
  }
};

class FileInterface : public nix::Application::Interface
{
public:
  PtrChunk create_chunk(nix::Dims3D dims, nix::Bool3D has_dim, int id) override
  {
    return std::make_unique<FileChunk>(dims, has_dim, id);
  }
};

class FileApplication : public GpuApplication
{
  std::string input;
  int         nrebalance = 0;

public:
  FileApplication(int argc, char** argv) : GpuApplication(argc, argv, std::make_shared<FileInterface>()) {}

  // the one hook Application::initialize leaves for the application (application.cpp:247-249)
  void initialize_domain() override
  {
    json opt     = cfgparser->get_application()["option"];
    input        = opt.value("input", std::string("."));
    order        = opt.value("order", 2);
    nb           = opt.value("nb", 2);
    strict_fp    = opt.value("strict", true);
    field_solver = opt.value("field_solver", false);
    cfj          = opt.value("cfj", 1.0);
    cc           = opt.value("cc", 1.0);
    for (auto& s : opt["species"]) species.push_back(SpeciesSpec{s[0].get<double>(), s[1].get<double>()});
  }

  bool rebalance() override
  {
    bool ran = GpuApplication::rebalance();
    nrebalance += ran ? 1 : 0;
    return ran;
  }

  void finalize() override
  {
    sync_host_all();
    long long total = 0;
    for (auto& c : chunkvec) {
      auto*       g  = static_cast<GpuChunk*>(c.get());
      std::string id = std::to_string(g->get_id());
      write_bin(input + "/out_uf_" + id + ".bin", g->uf.data(), g->uf.size());
      write_bin(input + "/out_uj_" + id + ".bin", g->uj.data(), g->uj.size());
      for (size_t is = 0; is < g->up.size(); is++) {
        write_bin(input + "/out_xu_" + id + "_" + std::to_string(is) + ".bin", g->up[is]->xu.data(), (size_t)g->up[is]->Np * 7);
        total += g->up[is]->Np;
      }
    }
    json summary = {{"steps", npush}, {"rebalance_calls_that_ran", nrebalance}, {"domain_builds", nrebuild},
                    {"particles", total}, {"launches", (long long)nixb200_launch_count()}, {"curstep", curstep}};
    std::ofstream(input + "/app_summary.json") << summary.dump(1) << std::endl;
    domain.reset(); // device memory goes before MPI does
    nix::Application::finalize();
  }
};

int main(int argc, char** argv)
{
  try {
    FileApplication app(argc, argv);
    return app.main();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "app_main failed: %s\n", e.what());
    return 1;
  }
}
