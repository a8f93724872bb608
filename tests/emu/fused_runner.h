// Runs a generated fused kernel on the CPU (TEST INFRASTRUCTURE): included after the generated
// source by the wrapper translation unit the test harness writes (tests/fused_emu_lib.py).
#pragma once
namespace {
struct FusedArgs {
  const pystl::TileProgram* P;
  const pystl::TileGeom* g;
  const float* signal;
  float* out;
  double* mean;
  const float* grad_out;
  float scale;
  float* grad_signal;
};
void fused_body(void* a) {
  FusedArgs* A = static_cast<FusedArgs*>(a);
#if defined(PF_EMU_BACKWARD)
  PF_EMU_KERNEL(*A->P, *A->g, A->signal, A->grad_out, A->scale, A->grad_signal, A->out, A->mean);
#else
  PF_EMU_KERNEL(*A->P, *A->g, A->signal, A->out, A->mean);
#endif
}
}  // namespace
extern "C" int pystl_fused_emu_run(const void* P, const void* g, const float* signal, float* out, double* mean,
                                   const float* grad_out, float scale, float* grad_signal, int smem_floats) {
  FusedArgs a{static_cast<const pystl::TileProgram*>(P), static_cast<const pystl::TileGeom*>(g), signal, out, mean, grad_out, scale, grad_signal};
  for (long long b = 0; b < a.g->B; ++b)
    for (int j = 0; j < a.g->n_tiles; ++j) simt::run_cta((int)b, j, a.g->nt, (size_t)smem_floats * 4, fused_body, &a);
  return 0;
}
