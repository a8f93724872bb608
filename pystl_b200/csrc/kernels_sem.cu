// One semantics' kernels (compiled with -DPYSTL_SEM=0..4, see Makefile).
#include <cstdarg>
#include <cstdio>

#include "kernels.cuh"

namespace pystl {

#ifndef PYSTL_SEM
#error "compile with -DPYSTL_SEM=<semantics id>"
#endif

static int set_err(char* err, int err_len, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err, err_len, fmt, ap);
  va_end(ap);
  return code;
}

template <int SEM>
int launch_fwd_sem(const Lowered& lw, const TileGeom& g, const float* signal, float* out, double* mean_acc,
                   cudaStream_t stream, char* err, int err_len) {
  const size_t smem = (size_t)g.smem_floats * 4;
  const int64_t grid = g.B * g.n_tiles;
  if (grid > 0x7fffffffLL) return set_err(err, err_len, PYSTL_E_UNSUPPORTED, "grid of %lld tiles exceeds 2^31-1", (long long)grid);
  cudaError_t e;
#define PYSTL_LAUNCH(MAXT, MINB)                                                                                  \
  do {                                                                                                            \
    auto kern = fwd_kernel<SEM, MAXT, MINB>;                                                                      \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                       \
    if (e != cudaSuccess) return set_err(err, err_len, PYSTL_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); \
    kern<<<(unsigned)grid, g.nt, smem, stream>>>(lw.prog, g, signal, out, mean_acc);                              \
  } while (0)
  if (g.nt <= 256) PYSTL_LAUNCH(256, PYSTL_MINB_256);
  else if (g.nt <= 512) PYSTL_LAUNCH(512, PYSTL_MINB_512);
  else PYSTL_LAUNCH(1024, 1);
#undef PYSTL_LAUNCH
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(err, err_len, PYSTL_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

// cumulative has no backward kernel yet
#if PYSTL_SEM != 2
template <int SEM>
int launch_bwd_sem(const Lowered& lw, const TileGeom& g, const float* signal, const float* grad_out, float scale,
                   float* grad_signal, cudaStream_t stream, char* err, int err_len) {
  const size_t smem = (size_t)g.smem_floats * 4;
  const int64_t grid = g.B * g.n_tiles;
  if (grid > 0x7fffffffLL) return set_err(err, err_len, PYSTL_E_UNSUPPORTED, "grid of %lld tiles exceeds 2^31-1", (long long)grid);
  cudaError_t e;
#define PYSTL_LAUNCH(MAXT, MINB)                                                                                  \
  do {                                                                                                            \
    auto kern = bwd_kernel<SEM, MAXT, MINB>;                                                                      \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                       \
    if (e != cudaSuccess) return set_err(err, err_len, PYSTL_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); \
    kern<<<(unsigned)grid, g.nt, smem, stream>>>(lw.prog, g, signal, grad_out, scale, grad_signal);               \
  } while (0)
  if (g.nt <= 256) PYSTL_LAUNCH(256, PYSTL_MINB_256);
  else if (g.nt <= 512) PYSTL_LAUNCH(512, PYSTL_MINB_512);
  else PYSTL_LAUNCH(1024, 1);
#undef PYSTL_LAUNCH
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(err, err_len, PYSTL_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}
template int launch_bwd_sem<PYSTL_SEM>(const Lowered&, const TileGeom&, const float*, const float*, float, float*,
                                       cudaStream_t, char*, int);
#endif

template int launch_fwd_sem<PYSTL_SEM>(const Lowered&, const TileGeom&, const float*, float*, double*, cudaStream_t,
                                       char*, int);

}  // namespace pystl
