// app_check.cpp -- compile check: GpuApplication / GpuInterface / GpuChunk against the reference's
// application.hpp and chunk.hpp (with the single-rank MPI stand-in of host/stub when MPI is absent).
#include "nixb200_host.hpp"

nix::Application* make_gpu_application(int argc, char** argv)
{
  return new nixb200host::GpuApplication(argc, argv);
}
