// libpystl_b200.so -- sm_100a kernels + the C ABI of include/pystl_b200.h.
//
// One CTA evaluates the whole formula over one tile of one trace
// (tile_eval.h); the signal tile is staged in shared memory by a TMA bulk copy
// (cp.async.bulk, falling back to vector loads when the tile is not 16-byte
// aligned), every intermediate node value stays in shared memory, only the
// root robustness goes back to HBM.  Bandwidth-bound windowed reductions: no
// tensor cores on purpose.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/pystl_b200.h"
#include "program.h"
#include "tile_eval.h"

namespace pystl {

// ------------------------------------------------------------------ device executor
struct DevExec {
  int tid_, nt_;
  float stash_;
  __device__ __forceinline__ DevExec() : tid_(threadIdx.x), nt_(blockDim.x), stash_(0.f) {}
  __device__ __forceinline__ int t0() const { return tid_; }
  __device__ __forceinline__ int t1() const { return tid_ + 1; }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void put(int, float v) { stash_ = v; }

  __device__ __forceinline__ float all_max(float* red) {
    float v = stash_;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = tid_ & 31, warp = tid_ >> 5, nw = nt_ >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    v = lane < nw ? red[lane] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
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

  // a[w*n + c] <- comb(a[c+1], ..., a[n-1]) in time order (identity for the last chunk); n == nt_
  template <int AW, class OP>
  __device__ __forceinline__ void excl_suffix_scan(float* a, int n, float* red) {
    using A = typename OP::A;
    const int lane = tid_ & 31, warp = tid_ >> 5, nw = nt_ >> 5;
    A v;
#pragma unroll
    for (int w = 0; w < AW; ++w) v.x[w] = a[w * n + tid_];
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
    for (int w = 0; w < AW; ++w) a[w * n + tid_] = excl.x[w];
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
template <int SEM, int MAXT>
__global__ void __launch_bounds__(MAXT) fwd_kernel(const __grid_constant__ TileProgram P, const __grid_constant__ TileGeom g,
                                                   const float* __restrict__ signal, float* __restrict__ out,
                                                   double* __restrict__ mean_acc) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bar;
  DevExec ex;
  if (ex.tid_ == 0) mbar_init(&bar, 1);
  __syncthreads();

  const int64_t tile = blockIdx.x;
  const int64_t b = tile / g.n_tiles;
  const int j = (int)(tile - b * g.n_tiles);
  Ctx c;
  c.smem = smem;
  c.g = g;
  c.P = &P;
  c.e0 = (int64_t)j * g.lout;
  c.T = g.T;
  c.rp = make_rp(P);

  float* raw = c.scr(0);
  const int nlive = c.n_live();
  stage_raw(raw, signal + (b * g.T + c.e0) * g.D, nlive * g.D, &bar, ex.tid_, ex.nt_);

  run_program<SEM>(ex, c, raw);
  constexpr int VW = Tr<SEM>::VW;
  store_root<SEM>(ex, c, out + b * g.T * VW, g.lout, mean_acc);
}

}  // namespace pystl

// ====================================================================== C ABI
using namespace pystl;

static thread_local char g_err[512] = "";

static int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static int forced_threads() {
  const char* e = getenv("PYSTL_B200_FORCE_THREADS");  // tests: multi-tile path on short traces
  return e ? atoi(e) : 0;
}

static int prepare(const pystl_formula_t* f, int64_t B, int64_t T, int32_t D, Lowered* lw, TileGeom* g) {
  int rc = lower_formula(reinterpret_cast<const ::pystl_formula*>(f), lw, g_err, sizeof(g_err));
  if (rc) return rc;
  return plan_geometry(*lw, B, T, D, forced_threads(), g, g_err, sizeof(g_err));
}

extern "C" int pystl_b200_abi_version(void) { return PYSTL_B200_ABI_VERSION; }
extern "C" const char* pystl_b200_last_error(void) { return g_err; }

extern "C" int pystl_b200_plan(const pystl_formula_t* f, int64_t T, int32_t D, int32_t backward, pystl_plan_t* plan) {
  if (!plan) return set_err(PYSTL_E_INVALID, "plan is NULL");
  if (backward) return set_err(PYSTL_E_UNSUPPORTED, "backward planning is not implemented yet");
  Lowered lw;
  TileGeom g;
  int rc = prepare(f, 1, T, D, &lw, &g);
  if (rc) return rc;
  plan->out_channels = lw.out_channels;
  plan->lookahead = lw.lookahead;
  plan->start_slack = lw.prog.start_slack;
  plan->raises_always = lw.prog.raises_always;
  plan->threads = g.nt;
  plan->tile_ext = g.lext;
  plan->tile_out = g.lout;
  plan->n_tiles = g.n_tiles;
  plan->smem_bytes = g.smem_floats * 4;
  plan->n_instr = lw.prog.n_instr;
  plan->n_planes = lw.prog.n_slots;
  return PYSTL_OK;
}

template <int SEM>
static int launch_fwd(const Lowered& lw, const TileGeom& g, const float* signal, float* out, double* mean_acc,
                      cudaStream_t stream) {
  const size_t smem = (size_t)g.smem_floats * 4;
  const int64_t grid = g.B * g.n_tiles;
  if (grid > 0x7fffffffLL) return set_err(PYSTL_E_UNSUPPORTED, "grid of %lld tiles exceeds 2^31-1", (long long)grid);
  cudaError_t e;
#define PYSTL_LAUNCH(MAXT)                                                                                       \
  do {                                                                                                           \
    auto kern = fwd_kernel<SEM, MAXT>;                                                                           \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                      \
    if (e != cudaSuccess) return set_err(PYSTL_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));       \
    kern<<<(unsigned)grid, g.nt, smem, stream>>>(lw.prog, g, signal, out, mean_acc);                             \
  } while (0)
  if (g.nt <= 256) PYSTL_LAUNCH(256);
  else if (g.nt <= 512) PYSTL_LAUNCH(512);
  else PYSTL_LAUNCH(1024);
#undef PYSTL_LAUNCH
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(PYSTL_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

extern "C" int pystl_b200_eval_forward(const pystl_formula_t* f, const float* signal, int64_t B, int64_t T, int32_t D,
                                       float* out, double* mean_acc, void* stream) {
  Lowered lw;
  TileGeom g;
  int rc = prepare(f, B, T, D, &lw, &g);
  if (rc) return rc;
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!signal || !out) return set_err(PYSTL_E_INVALID, "signal / out is NULL");
  if (g.nt % 32 != 0) return set_err(PYSTL_E_INVALID, "CTA size %d is not a multiple of 32", g.nt);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (lw.prog.sem) {
    case SEM_CLASSIC: return launch_fwd<SEM_CLASSIC>(lw, g, signal, out, mean_acc, s);
    case SEM_SMOOTH: return launch_fwd<SEM_SMOOTH>(lw, g, signal, out, mean_acc, s);
    case SEM_CUMULATIVE: return launch_fwd<SEM_CUMULATIVE>(lw, g, signal, out, mean_acc, s);
    case SEM_AGM: return launch_fwd<SEM_AGM>(lw, g, signal, out, mean_acc, s);
    case SEM_DGMSR: return launch_fwd<SEM_DGMSR>(lw, g, signal, out, mean_acc, s);
    default: return set_err(PYSTL_E_INVALID, "unknown semantics");
  }
}

extern "C" int pystl_b200_eval_backward(const pystl_formula_t*, const float*, int64_t, int64_t, int32_t, const float*,
                                        float, float*, void*) {
  return set_err(PYSTL_E_UNSUPPORTED, "backward kernels are not built yet");
}
