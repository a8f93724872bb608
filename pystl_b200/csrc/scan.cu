// Open-ended windows on traces of any length: G[a, inf) / F[a, inf) as a SUFFIX SCAN of a value trace.
//
// The reference clips an interval without an end at the horizon (api.py:49-54: end None -> horizon - 1) and hands the
// window [t + a, T - 1] to the semantics' temporal_always / temporal_eventually (classic.py:39-57, smooth.py:47-65,
// agm.py:92-110, dgmsr.py:58-86).  Inside one tile the formula kernels evaluate that as a bounded window (interpreter
// tier: T_WIN_SUFFIX).  Beyond one tile (T > 8192) the value at t depends on everything up to the horizon, so it is
// computed here instead, on the value trace y[b, t] of the operand (itself produced by the formula kernels):
//
//     S[u]   = (+)_{k = u}^{T-1} lift(y_k)                       one suffix scan per trace
//     out[t] = fin(S[u], T - u),   u = min(t + a, T - 1)          (t + a > T - 1: the reference raises; those outputs hold the
//                                                                  last valid value and are masked by the caller's start_slack)
//
// and the adjoint is the mirrored PREFIX scan of per-output coefficients (tile_bwd.h: Adj<SEM, MODE>).  The host side
// (pystl_b200/compiler.py: OpenLeaf) feeds `out` back to the formula kernels as a leaf channel.
//
// One WARP per trace: every lane owns a contiguous segment of the trace, reduces it, the 32 segment totals are scanned
// with shuffles, and a second walk over the segment writes the outputs.  The trace is read twice and written once.
// Smooth carries (max exponent, scaled sum) pairs, so no trace-wide shift is needed and the result is exact for any
// temperature.  Classic records the arg-min / arg-max (first index on ties, like torch.min / max(dim)) for the gradient.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/pystl_b200.h"
#include "tile_bwd.h"

namespace pystl {
namespace {

constexpr int kWarpsPerBlock = 4;

// ---------------------------------------------------------------- forward reducers: ident / lift / comb(earlier, later) / fin
template <int SEM, int MODE>
struct ScanRed {  // agm, dgmsr: the reducers of the formula kernels
  using R = Red<SEM, MODE>;
  using A = typename R::A;
  static __device__ A ident() { return R::ident(); }
  static __device__ A lift(float y, int, const RP& rp) { return R::lift(Vec<1>{{y}}, rp, 1.f); }
  static __device__ A comb(A a, A b) { return R::comb(a, b); }
  static __device__ float fin(A a, float n, const RP& rp) { return R::fin(a, n, rp).x[0]; }
  static __device__ int arg(A) { return 0; }
};
template <int MODE>
struct ScanRed<SEM_CLASSIC, MODE> {  // (value, position): the earlier operand keeps ties
  using A = Vec<2>;
  static __device__ A ident() { return A{{MODE == MODE_AND ? f_inf() : -f_inf(), i2f(0)}}; }
  static __device__ A lift(float y, int k, const RP&) { return A{{y, i2f(k)}}; }
  static __device__ A comb(A a, A b) {
    const bool tb = MODE == MODE_AND ? (b.x[0] < a.x[0]) : (b.x[0] > a.x[0]);
    return tb ? b : a;
  }
  static __device__ float fin(A a, float, const RP&) { return a.x[0]; }
  static __device__ int arg(A a) { return f2i(a.x[1]); }
};
template <int MODE>
struct ScanRed<SEM_SMOOTH, MODE> {  // (m, s): sum of 2^e = s * 2^m, m the running maximum of the exponents e = -+ k y
  using A = Vec<2>;
  static __device__ A ident() { return A{{-f_inf(), 0.f}}; }
  static __device__ A lift(float y, int, const RP& rp) { return A{{(MODE == MODE_AND ? -y : y) * rp.k, 1.f}}; }
  static __device__ A comb(A a, A b) {
    const float m = fmaxf(a.x[0], b.x[0]);
    if (m == -f_inf()) return A{{m, 0.f}};
    return A{{m, a.x[1] * fast_ex2(a.x[0] - m) + b.x[1] * fast_ex2(b.x[0] - m)}};
  }
  static __device__ float fin(A a, float, const RP& rp) {
    const float l = (a.x[0] + fast_lg2(a.x[1])) * rp.invk;
    return MODE == MODE_AND ? -l : l;
  }
  static __device__ int arg(A) { return 0; }
};

template <class A>
__device__ A shfl_down_acc(A v, int d) {
  A r;
#pragma unroll
  for (int w = 0; w < (int)(sizeof(A) / sizeof(float)); ++w) r.x[w] = __shfl_down_sync(0xffffffffu, v.x[w], d);
  return r;
}
template <class A>
__device__ A shfl_up_acc(A v, int d) {
  A r;
#pragma unroll
  for (int w = 0; w < (int)(sizeof(A) / sizeof(float)); ++w) r.x[w] = __shfl_up_sync(0xffffffffu, v.x[w], d);
  return r;
}

template <int SEM, int MODE>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) open_window_fwd(const float* __restrict__ y, float* __restrict__ out,
                                                                       int* __restrict__ arg, long long B, int T, int a, RP rp) {
  using S = ScanRed<SEM, MODE>;
  using A = typename S::A;
  const long long b = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const float* yb = y + b * T;
  float* ob = out + b * T;
  int* ab = arg ? arg + b * T : nullptr;
  const int seg = (T + 31) / 32;
  const int lo = lane * seg < T ? lane * seg : T, hi = lo + seg < T ? lo + seg : T;
  // pass 1: the segment's total
  A tot = S::ident();
  for (int k = hi - 1; k >= lo; --k) tot = S::comb(S::lift(yb[k], k, rp), tot);
  // inclusive suffix scan over the lanes, then exclusive
  A inc = tot;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const A other = shfl_down_acc(inc, d);
    if (lane + d < 32) inc = S::comb(inc, other);
  }
  A run = shfl_down_acc(inc, 1);
  if (lane == 31) run = S::ident();
  // pass 2: S[u] for the segment's u, stored at t = u - a; the last a outputs repeat S[T - 1]'s value
  for (int u = hi - 1; u >= lo; --u) {
    run = S::comb(S::lift(yb[u], u, rp), run);
    const float v = S::fin(run, (float)(T - u), rp);
    const int t = u - a;
    if (t >= 0) {
      ob[t] = v;
      if (ab) ab[t] = S::arg(run);
    }
    if (u == T - 1) {  // t + a > T - 1: the reference raises (masked by the caller); keep a finite value of the right magnitude
      for (int tt = (T - a > 0 ? T - a : 0); tt < T; ++tt) {
        if (tt == t) continue;
        ob[tt] = v;
        if (ab) ab[tt] = -1;
      }
    }
  }
}

// ---------------------------------------------------------------- backward
// grad_y[k] = in_factor( sum_{t <= k - a, t valid} out_coeff(g_t, out_t, T - t - a), y_k ): a prefix scan of coefficients
template <int SEM, int MODE>
struct ScanAdj {  // agm, dgmsr
  using AD = Adj<SEM, MODE>;
  using A = Vec<2>;
  static __device__ A ident() { return A{{0.f, 0.f}}; }
  static __device__ A coeff(float g, float o, float n, const RP& rp) {
    bool lost;
    return g != 0.f ? AD::out_coeff(g, o, n, rp, lost) : ident();
  }
  static __device__ A comb(A a, A b) { return A{{a.x[0] + b.x[0], a.x[1] + b.x[1]}}; }
  static __device__ float apply(A w, float y, const RP& rp) { return AD::in_factor(w, y, 1.f, rp); }
};
template <int MODE>
struct ScanAdj<SEM_SMOOTH, MODE> {  // d out_t / d y_k = 2^(k (out_t - y_k)) (AND), 2^(k (y_k - out_t)) (OR): <= 1, exact
  using A = Vec<2>;                 // (m, s): sum of g_t 2^(e_t) = s * 2^m, e_t = +- k out_t
  static __device__ A ident() { return A{{-f_inf(), 0.f}}; }
  static __device__ A coeff(float g, float o, float, const RP& rp) {
    return g != 0.f ? A{{(MODE == MODE_AND ? o : -o) * rp.k, g}} : ident();
  }
  static __device__ A comb(A a, A b) {
    const float m = fmaxf(a.x[0], b.x[0]);
    if (m == -f_inf()) return A{{m, 0.f}};
    return A{{m, a.x[1] * fast_ex2(a.x[0] - m) + b.x[1] * fast_ex2(b.x[0] - m)}};
  }
  static __device__ float apply(A w, float y, const RP& rp) {
    if (w.x[0] == -f_inf()) return 0.f;
    return w.x[1] * fast_ex2(w.x[0] + (MODE == MODE_AND ? -y : y) * rp.k);
  }
};

template <int SEM, int MODE>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) open_window_bwd(const float* __restrict__ y, const float* __restrict__ out,
                                                                       const float* __restrict__ grad_out, float* __restrict__ grad_y,
                                                                       long long B, int T, int a, RP rp) {
  using S = ScanAdj<SEM, MODE>;
  using A = typename S::A;
  const long long b = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const float* yb = y + b * T;
  const float* ob = out + b * T;
  const float* gb = grad_out + b * T;
  float* dy = grad_y + b * T;
  const int nvalid = T - a;  // outputs t < nvalid exist
  const int seg = (T + 31) / 32;
  const int lo = lane * seg < T ? lane * seg : T, hi = lo + seg < T ? lo + seg : T;
  auto coeff = [&](int t) { return t < nvalid ? S::coeff(gb[t], ob[t], (float)(T - t - a), rp) : S::ident(); };
  A tot = S::ident();
  for (int t = lo; t < hi; ++t) tot = S::comb(tot, coeff(t));
  A inc = tot;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const A other = shfl_up_acc(inc, d);
    if (lane >= d) inc = S::comb(other, inc);
  }
  A run = shfl_up_acc(inc, 1);
  if (lane == 0) run = S::ident();
  // P[t] for the segment's t serves grad_y[t + a]; samples k < a belong to no window
  for (int t = lo; t < hi; ++t) {
    run = S::comb(run, coeff(t));
    const int k = t + a;
    if (k < T) dy[k] = S::apply(run, yb[k], rp);
    if (t < a) dy[t] = 0.f;
  }
}
// classic: every valid output routes its cotangent to the recorded arg-min / arg-max
__global__ void open_window_bwd_classic(const int* __restrict__ arg, const float* __restrict__ grad_out, float* __restrict__ grad_y,
                                        long long B, int T, int a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int t = (int)(i % T);
  if (t >= T - a) return;
  const float g = grad_out[i];
  const int k = arg[i];
  if (g != 0.f && k >= 0) atomicAdd(&grad_y[(i / T) * T + k], g);
}

thread_local char* g_err = nullptr;  // the caller's message buffer (pystl_b200.cu: pystl_b200_last_error)
thread_local size_t g_err_cap = 0;
int scan_fail(int code, const char* fmt, ...) {
  if (!g_err || !g_err_cap) return code;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, g_err_cap, fmt, ap);
  va_end(ap);
  return code;
}

RP make_rp(float temperature, float eps, int p) {
  RP r{};
  const double tau = temperature > 0.f ? temperature : 1.0;
  const int pp = p < 1 ? 1 : p;
  const double e = eps > 0.f ? eps : 1e-8;
  r.k = (float)(1.4426950408889634 / tau);
  r.invk = (float)(tau * 0.6931471805599453);
  r.shift = 0.f;
  r.eps = (float)e;
  r.sqrt_eps = (float)sqrt(e);
  r.ln_eps = (float)log(e);
  r.eps_p = (float)pow(e, (double)pp);
  r.inv2p = (float)(0.5 / pp);
  r.p = pp;
  return r;
}

int check(int semantics, int mode, int a, long long B, long long T, float temperature, float eps, int p) {
  if (semantics != SEM_CLASSIC && semantics != SEM_SMOOTH && semantics != SEM_AGM && semantics != SEM_DGMSR)
    return scan_fail(PYSTL_E_UNSUPPORTED, "open-ended windows beyond one tile: classical, smooth, agm and dgmsr (cumulative values are pairs)");
  if (mode != MODE_AND && mode != MODE_OR) return scan_fail(PYSTL_E_INVALID, "mode must be 0 (always) or 1 (eventually)");
  if (a < 0) return scan_fail(PYSTL_E_INVALID, "Interval start must be >= 0.");
  if (B < 0 || T < 0 || T >= (1ll << 31) - 65536) return scan_fail(PYSTL_E_INVALID, "bad shape");
  if (semantics == SEM_SMOOTH && !(temperature > 0.f)) return scan_fail(PYSTL_E_INVALID, "temperature must be > 0.");
  if (semantics == SEM_DGMSR && (!(eps > 0.f) || p < 1)) return scan_fail(PYSTL_E_INVALID, "dgmsr needs eps > 0 and p >= 1");
  return PYSTL_OK;
}

}  // namespace

#define PYSTL_SCAN_DISPATCH(KERNEL, ...)                                               \
  do {                                                                                 \
    const unsigned grid = (unsigned)((B + kWarpsPerBlock - 1) / kWarpsPerBlock);       \
    const dim3 block(32 * kWarpsPerBlock);                                             \
    if (semantics == SEM_CLASSIC) {                                                    \
      if (mode == MODE_AND) KERNEL<SEM_CLASSIC, MODE_AND><<<grid, block, 0, st>>>(__VA_ARGS__); \
      else KERNEL<SEM_CLASSIC, MODE_OR><<<grid, block, 0, st>>>(__VA_ARGS__);          \
    } else if (semantics == SEM_SMOOTH) {                                              \
      if (mode == MODE_AND) KERNEL<SEM_SMOOTH, MODE_AND><<<grid, block, 0, st>>>(__VA_ARGS__); \
      else KERNEL<SEM_SMOOTH, MODE_OR><<<grid, block, 0, st>>>(__VA_ARGS__);           \
    } else if (semantics == SEM_AGM) {                                                 \
      if (mode == MODE_AND) KERNEL<SEM_AGM, MODE_AND><<<grid, block, 0, st>>>(__VA_ARGS__); \
      else KERNEL<SEM_AGM, MODE_OR><<<grid, block, 0, st>>>(__VA_ARGS__);              \
    } else {                                                                           \
      if (mode == MODE_AND) KERNEL<SEM_DGMSR, MODE_AND><<<grid, block, 0, st>>>(__VA_ARGS__); \
      else KERNEL<SEM_DGMSR, MODE_OR><<<grid, block, 0, st>>>(__VA_ARGS__);            \
    }                                                                                  \
  } while (0)

int open_window_forward(int32_t semantics, int32_t mode, int32_t a, float temperature, float eps, int32_t p, const float* y,
                        int64_t B, int64_t T, float* out, int32_t* arg, void* stream, char* err, size_t err_cap) {
  g_err = err;
  g_err_cap = err_cap;
  int rc = check(semantics, mode, a, B, T, temperature, eps, p);
  if (rc) return rc;
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!y || !out) return scan_fail(PYSTL_E_INVALID, "null buffer");
  if (semantics == SEM_CLASSIC && !arg) return scan_fail(PYSTL_E_INVALID, "classical needs the arg buffer");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return scan_fail(PYSTL_E_CUDA, "no CUDA device: the b200 backend has no CPU path");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const RP rp = make_rp(temperature, eps, p);
  int* argp = semantics == SEM_CLASSIC ? arg : nullptr;
  PYSTL_SCAN_DISPATCH(open_window_fwd, y, out, argp, (long long)B, (int)T, (int)a, rp);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return scan_fail(PYSTL_E_CUDA, "open_window_fwd: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

int open_window_backward(int32_t semantics, int32_t mode, int32_t a, float temperature, float eps, int32_t p, const float* y,
                         const float* out, const int32_t* arg, const float* grad_out, int64_t B, int64_t T, float* grad_y,
                         void* stream, char* err, size_t err_cap) {
  g_err = err;
  g_err_cap = err_cap;
  int rc = check(semantics, mode, a, B, T, temperature, eps, p);
  if (rc) return rc;
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!y || !out || !grad_out || !grad_y) return scan_fail(PYSTL_E_INVALID, "null buffer");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return scan_fail(PYSTL_E_CUDA, "no CUDA device: the b200 backend has no CPU path");
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (semantics == SEM_CLASSIC) {
    if (!arg) return scan_fail(PYSTL_E_INVALID, "classical needs the arg buffer");
    cudaError_t e = cudaMemsetAsync(grad_y, 0, (size_t)B * (size_t)T * sizeof(float), st);
    if (e != cudaSuccess) return scan_fail(PYSTL_E_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const long long n = (long long)B * T;
    open_window_bwd_classic<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(arg, grad_out, grad_y, (long long)B, (int)T, (int)a);
  } else {
    const RP rp = make_rp(temperature, eps, p);
    const unsigned grid = (unsigned)((B + kWarpsPerBlock - 1) / kWarpsPerBlock);
    const dim3 block(32 * kWarpsPerBlock);
#define PYSTL_BWD(SEM)                                                                                                        \
  do {                                                                                                                        \
    if (mode == MODE_AND) open_window_bwd<SEM, MODE_AND><<<grid, block, 0, st>>>(y, out, grad_out, grad_y, (long long)B, (int)T, (int)a, rp); \
    else open_window_bwd<SEM, MODE_OR><<<grid, block, 0, st>>>(y, out, grad_out, grad_y, (long long)B, (int)T, (int)a, rp);   \
  } while (0)
    if (semantics == SEM_SMOOTH) PYSTL_BWD(SEM_SMOOTH);
    else if (semantics == SEM_AGM) PYSTL_BWD(SEM_AGM);
    else PYSTL_BWD(SEM_DGMSR);
#undef PYSTL_BWD
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return scan_fail(PYSTL_E_CUDA, "open_window_bwd: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

}  // namespace pystl
