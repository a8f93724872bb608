// Host side of the fused tier: code generation (codegen.cpp), tile geometry, and the run-time
// compiler + cubin cache (jit.cpp).
#pragma once
#include <stdint.h>

#include <string>

#include "program.h"

namespace pystl {

struct FusedSource {
  std::string source;  // CUDA source of the kernel (includes fused_rt.cuh)
  std::string name;    // its extern "C" name: pystl_fused_<pass>_<semantics>
  int smem_floats = 0;
  int nt = 0;
  int minb = 1;
};

// Tile geometry of the fused tier for (formula, T, D).  PYSTL_E_UNSUPPORTED when the formula has
// no bounded look-ahead (the interpreter tier handles those).
int fused_geometry(const Lowered& lw, int64_t B, int64_t T, int32_t D, int force_nt, bool backward, TileGeom* g, char* msg,
                   int msg_len);

// Emit the kernel for a lowered (SSA) program.  PYSTL_E_UNSUPPORTED when the program uses an
// instruction the fused tier does not implement.
int fused_generate(const Lowered& lw, int D, int nt, int minb, bool backward, FusedSource* out, char* msg, int msg_len);

// ---- jit.cpp (CUDA side; not part of the emulator build)
struct FusedKernel {
  void* kernel = nullptr;  // cudaKernel_t
  int regs = 0;
};
// Compile (or fetch from the in-memory / on-disk cache) and load the kernel.
int fused_jit_get(const FusedSource& src, FusedKernel* out, char* msg, int msg_len);
// Compile into the on-disk cache only: needs no GPU (used at build time).
int fused_jit_precompile(const FusedSource& src, char* msg, int msg_len);
const char* fused_cache_dir();

}  // namespace pystl
