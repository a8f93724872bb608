// Device side of libpystl_b200.so: the CTA executor (one CUDA thread per lane), the TMA
// staging of the signal tile and the fused-formula kernel template.  Included by
// kernels_sem.cu, which instantiates it once per semantics (one object file each, so the
// library builds in parallel).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/pystl_b200.h"
#include "program.h"
#include "tile_eval.h"
#include "tile_bwd.h"

namespace pystl {

// ------------------------------------------------------------------ device executor
struct DevExec {
  int tid_, nt_;
  float stash_, stash2_;
  __device__ __forceinline__ DevExec() : tid_(threadIdx.x), nt_(blockDim.x), stash_(0.f), stash2_(0.f) {}
  __device__ __forceinline__ int t0() const { return tid_; }
  __device__ __forceinline__ int t1() const { return tid_ + 1; }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void put(int, float v) { stash_ = v; }
  __device__ __forceinline__ float get(int) const { return stash_; }
  __device__ __forceinline__ void atomic_add(float* p, float v) { atomicAdd(p, v); }

  // max over a warp with one REDUX: floats are mapped to integers of the same order
  static __device__ __forceinline__ float warp_max(float v) {
    int k = __float_as_int(v);
    k = k >= 0 ? k : k ^ 0x7fffffff;
    k = __reduce_max_sync(0xffffffffu, k);
    k = k >= 0 ? k : k ^ 0x7fffffff;
    return __int_as_float(k);
  }
  __device__ __forceinline__ void put2(int, float a, float b) {
    stash_ = a;
    stash2_ = b;
  }
  // two block-wide maxima at once
  __device__ __forceinline__ void all_max2(float* red, float& a, float& b) {
    float va = warp_max(stash_), vb = warp_max(stash2_);
    const int lane = tid_ & 31, warp = tid_ >> 5, nw = nt_ >> 5;
    if (lane == 0) {
      red[64 + warp] = va;
      red[96 + warp] = vb;
    }
    __syncthreads();
    a = warp_max(lane < nw ? red[64 + lane] : -INFINITY);
    b = warp_max(lane < nw ? red[96 + lane] : -INFINITY);
    __syncthreads();
  }
  __device__ __forceinline__ float all_max(float* red) {
    float v = warp_max(stash_);
    const int lane = tid_ & 31, warp = tid_ >> 5, nw = nt_ >> 5;
    if (lane == 0) red[64 + warp] = v;
    __syncthreads();
    v = warp_max(lane < nw ? red[64 + lane] : -INFINITY);
    __syncthreads();
    return v;
  }

  __device__ __forceinline__ void add_block_sum(float* red, double* acc) {
    float v = stash_;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = tid_ & 31, warp = tid_ >> 5, nw = nt_ >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
      double d = lane < nw ? (double)red[lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (lane == 0) atomicAdd(acc, d);
    }
    __syncthreads();
  }

  // a[w*stride + c] <- comb(a[c+1], ..., a[n-1]) in time order (identity for the last chunk); n == nt_
  template <int AW, class OP>
  __device__ __forceinline__ void excl_suffix_scan(float* a, int n, int stride, float* red) {
    (void)n;
    using A = typename OP::A;
    const int lane = tid_ & 31, warp = tid_ >> 5, nw = nt_ >> 5;
    A v;
#pragma unroll
    for (int w = 0; w < AW; ++w) v.x[w] = a[w * stride + tid_];
    // inclusive suffix scan inside the warp: v_lane = x_lane (+) ... (+) x_31
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      A other;
#pragma unroll
      for (int w = 0; w < AW; ++w) other.x[w] = __shfl_down_sync(0xffffffffu, v.x[w], o);
      if (lane + o < 32) v = OP::comb(v, other);
    }
    if (lane == 0) {
#pragma unroll
      for (int w = 0; w < AW; ++w) red[w * 32 + warp] = v.x[w];
    }
    A excl;
#pragma unroll
    for (int w = 0; w < AW; ++w) excl.x[w] = __shfl_down_sync(0xffffffffu, v.x[w], 1);
    if (lane == 31) excl = OP::ident();
    __syncthreads();
    A tail = OP::ident();
    for (int ww = nw - 1; ww > warp; --ww) {
      A t;
#pragma unroll
      for (int w = 0; w < AW; ++w) t.x[w] = red[w * 32 + ww];
      tail = OP::comb(t, tail);
    }
    excl = OP::comb(excl, tail);
#pragma unroll
    for (int w = 0; w < AW; ++w) a[w * stride + tid_] = excl.x[w];
    __syncthreads();
  }
};

// ------------------------------------------------------------------ TMA bulk load (1-D)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Stage `n` floats starting at `src` into shared `dst`.
__device__ __forceinline__ void stage_raw(float* dst, const float* __restrict__ src, int n, uint64_t* bar, int tid, int nt) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  const int n16 = aligned ? (n & ~3) : 0;
  if (n16 > 0) {
    if (tid == 0) {
      mbar_expect_tx(bar, (uint32_t)n16 * 4u);
      tma_load_1d(dst, src, (uint32_t)n16 * 4u, bar);
    }
  }
  for (int i = n16 + tid; i < n; i += nt) dst[i] = __ldg(src + i);  // unaligned trace or the <16-byte tail
  if (n16 > 0) mbar_wait(bar, 0);
  __syncthreads();
}

// ------------------------------------------------------------------ forward kernel
template <int SEM, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) fwd_kernel(const __grid_constant__ TileProgram P, const __grid_constant__ TileGeom g,
                                                   const float* __restrict__ signal, float* __restrict__ out,
                                                   double* __restrict__ mean_acc) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bar;
  DevExec ex;
  if (ex.tid_ == 0) mbar_init(&bar, 1);
  __syncthreads();

  const unsigned tile = blockIdx.x;
  const unsigned bu = tile / (unsigned)g.n_tiles;
  const int j = (int)(tile - bu * (unsigned)g.n_tiles);
  const int64_t b = bu;
  const int64_t e0 = (int64_t)j * g.lout;
  const Ctx c = make_ctx(smem, g, P, e0);

  float* raw = c.raw();
  stage_raw(raw, signal + (b * g.T + e0) * g.D, c.nlive * g.D, &bar, ex.tid_, ex.nt_);

  run_program<SEM>(ex, c, P, raw);
  constexpr int VW = Tr<SEM>::VW;
  store_root<SEM>(ex, c, P, out + (b * g.T + e0) * VW, g.lout, mean_acc);
}

// ------------------------------------------------------------------ backward kernel
// grad_out == nullptr: constant cotangent `scale` (the mean-robustness loss)
template <int SEM, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) bwd_kernel(const __grid_constant__ TileProgram P, const __grid_constant__ TileGeom g,
                                                        const float* __restrict__ signal, const float* __restrict__ grad_out,
                                                        float scale, float* __restrict__ grad_signal) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bar;
  DevExec ex;
  if (ex.tid_ == 0) mbar_init(&bar, 1);
  __syncthreads();

  const unsigned tile = blockIdx.x;
  const unsigned bu = tile / (unsigned)g.n_tiles;
  const int j = (int)(tile - bu * (unsigned)g.n_tiles);
  const int64_t b = bu;
  const int64_t t0 = (int64_t)j * g.lout;
  const int64_t e0 = t0 > g.hh ? t0 - g.hh : 0;
  const Ctx c = make_ctx(smem, g, P, e0);
  const int p_lo = (int)(t0 - e0);
  int p_hi = p_lo + g.lout;
  if (p_hi > c.nlive) p_hi = c.nlive;

  float* raw = c.raw();
  stage_raw(raw, signal + (b * g.T + e0) * g.D, c.nlive * g.D, &bar, ex.tid_, ex.nt_);
  run_backward<SEM>(ex, c, P, raw, grad_out ? grad_out + (b * g.T + e0) : nullptr, scale,
                    grad_signal + (b * g.T + e0) * g.D, p_lo, p_hi);
}

}  // namespace pystl
