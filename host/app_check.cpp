// app_check.cpp -- compile check: GpuApplication / GpuInterface / GpuChunk against the reference's
// application.hpp and chunk.hpp (with the single-rank MPI stand-in of host/stub when MPI is absent).
#include "nixb200_host.hpp"

nix::Application* make_gpu_application(int argc, char** argv)
{
  return new nixb200host::GpuApplication(argc, argv);
}

// ... and the device-side counterpart of XtensorPacker3D in the shape a ChunkWriter Packer functor has
// (diag/chunk_writer.hpp:118-146: size = packer(chunk, nullptr, 0); address = packer(chunk, buffer, address))
struct FieldPackerCheck {
  nixb200host::GpuPacker3D packer;
  size_t operator()(nixb200host::GpuChunk* chunk, uint8_t* buffer, int address)
  {
    address = (int)packer.pack_field(*chunk, 2, buffer, address);
    address = (int)packer.pack_moment(*chunk, 0, 2, buffer, address);
    return packer.pack_tracer(*chunk, 0, buffer, address);
  }
};

size_t packer_check(nixb200host::GpuChunk* chunk, uint8_t* buffer, int address)
{
  return FieldPackerCheck()(chunk, buffer, address);
}
