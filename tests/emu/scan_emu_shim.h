// Host stand-ins for the CUDA vocabulary of pystl_b200/csrc/scan.cu (TEST INFRASTRUCTURE, never part of the product): included by
// scan.cu when PYSTL_SCAN_EMU is defined, after tests/emu/simt_host.h.  Every CUDA thread is a fiber of simt_host.h; a `__shared__`
// array is a function-local static (one CTA runs at a time, all its fibers see the same storage); warp shuffles exchange through
// the fiber scheduler's per-thread slots, with the hardware's rule that a source lane outside the warp returns the caller's value.
#pragma once
#include <string.h>

struct ScanDim { unsigned x; };
namespace scan_emu {
inline ScanDim tidx() { return ScanDim{(unsigned)simt::tid()}; }
inline ScanDim bidx() { return ScanDim{(unsigned)simt::bid_x()}; }
inline unsigned& grid_dim() { static thread_local unsigned g = 1; return g; }
inline ScanDim gdim() { return ScanDim{grid_dim()}; }
inline float as_f(int v) { float f; memcpy(&f, &v, 4); return f; }
inline int as_i(float f) { int v; memcpy(&v, &f, 4); return v; }
}  // namespace scan_emu
#define threadIdx (scan_emu::tidx())
#define blockIdx (scan_emu::bidx())
#define gridDim (scan_emu::gdim())
#define __grid_constant__
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(n)
#define __restrict__
#define __shared__ static
#define __syncwarp() simt::syncwarp()
inline float __int_as_float(int u) { return scan_emu::as_f(u); }
inline float __shfl_sync(unsigned, float v, int src) { return simt::shfl_idx(v, src); }
inline int __shfl_sync(unsigned, int v, int src) { return scan_emu::as_i(simt::shfl_idx(scan_emu::as_f(v), src)); }
inline float __shfl_down_sync(unsigned, float v, int d) { return simt::shfl_down(v, d); }
inline float __shfl_up_sync(unsigned, float v, int d) {
  const int lane = simt::tid() & 31;
  return simt::shfl_idx(v, lane - d >= 0 ? lane - d : lane);
}
inline int __shfl_up_sync(unsigned, int v, int d) { return scan_emu::as_i(__shfl_up_sync(0u, scan_emu::as_f(v), d)); }
