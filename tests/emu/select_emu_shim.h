// Host stand-ins for the CUDA vocabulary of pystl_b200/csrc/select.cu (TEST INFRASTRUCTURE, never part of the product):
// included by select.cu when PYSTL_SELECT_EMU is defined, after tests/emu/simt_host.h.  Every CUDA thread is a fiber of
// simt_host.h; threadIdx / blockIdx / gridDim read the running fiber's coordinates.
#pragma once
#include <string.h>

struct float4 { float x, y, z, w; };
struct SelDim { unsigned x; };
namespace sel_emu {
inline unsigned& grid_dim() { static thread_local unsigned g = 1; return g; }
inline int& shared_int() { static thread_local int v = 0; return v; }  // the one __shared__ scalar of a kernel (a CTA at a time)
inline SelDim tidx() { return SelDim{(unsigned)simt::tid()}; }
inline SelDim bidx() { return SelDim{(unsigned)simt::bid_x()}; }
inline SelDim gdim() { return SelDim{grid_dim()}; }
}  // namespace sel_emu
#define threadIdx (sel_emu::tidx())
#define blockIdx (sel_emu::bidx())
#define gridDim (sel_emu::gdim())
#define __syncthreads() simt::sync()
#define __device__
#define __forceinline__ inline
#define __restrict__
#define SEL_KERNEL(NT) void
#define SEL_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(simt::dyn_smem())
#define SEL_SHARED(T, name) T& name = sel_emu::shared_int()
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
