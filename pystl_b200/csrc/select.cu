// CT-STL on value traces: the cumulative-time operator C^tau_I (k-th largest of a window) and the Until with an EXCLUSIVE
// left prefix -- the two operators in which CtstlSemantics (ctstl.py:38-127) differs from the classical semantics.
//
//   C^tau_I:  out[b, t] = k-th largest of y[b, t + a .. min(t + b_end, T - 1)],  k = ceil(tau / delta)      (ctstl.py:18-35, 107-127)
//             -inf when the clipped window holds fewer than k samples (ctstl.py:118-125; an empty window too); NaN when the
//             window holds a NaN (an operand value the reference raises on)
//   Until:    out[b, t] = max_{i = a .. min(b_end, T - 1 - t)} min( r[t + i], min_{k < i} l[t + k] )         (ctstl.py:78-105)
//             (the prefix of candidate i stops BEFORE t + i; candidate 0 is r[t] alone); NaN where the window is empty (the
//             reference raises ValueError there)
//
// A selection is not a scan: there is no associative summary of a window from which the k-th largest of a shifted window
// follows, so every output selects from its own window.  The kernel does that by BISECTION ON THE KEY: floats are mapped
// ordered by an order-preserving 32-bit key, the tile (outputs + window halo) is staged into shared memory with coalesced
// loads, and the answer's key is built bit by bit from the top -- 32 counting passes `#(sample >= candidate) >= k` over the
// window, the candidate key turned back into a float so that the compare is an FSET and the count an FADD (two pipes).
// A thread owns 4 consecutive outputs whose windows overlap in all but 3 samples: every 128-bit shared-memory read feeds
// 16 compare-and-count operations (no bank conflicts: consecutive threads read consecutive 16-byte groups).  The work is
// 32 (W + 3) compares per output whatever k is; the samples are read from HBM once and the result written once.
//
// The gradient goes to ONE sample per output (the selected one; among equal samples the one a stable ascending sort puts at
// position n - k, i.e. occurrence number #(key >= answer) - k in window order -- what jnp.sort's gather does in the
// reference's jax twin, jax/semantics.py:378-386).  Backward is a GATHER over the recorded indices: sample s collects from
// the outputs t in [s - b_end, s - a] whose selection it was -- no atomics, deterministic.
#if defined(PYSTL_SELECT_EMU)
// Host build for the CPU test tier (tests/emu/select_emu.cpp): the kernels below are compiled with g++, every CUDA thread a fiber
// of tests/emu/simt_host.h (barriers are cooperative yields).  Test infrastructure; the product is the CUDA build.
#include "select_emu_shim.h"
#else
#include <cuda_runtime.h>
#define SEL_KERNEL(NT) __global__ void __launch_bounds__(NT)
#define SEL_DYN_SMEM(T, name) extern __shared__ __align__(16) T name[]
#define SEL_SHARED(T, name) __shared__ T name
#endif

#include <cstdarg>
#include <cstdio>

#include "../../include/pystl_b200.h"

namespace pystl {
namespace {

constexpr int kSelThreads = 128;           // threads per CTA
constexpr int kSelOut = 4 * kSelThreads;   // outputs per CTA
constexpr int kSelMaxWindow = 48 * 1024;   // samples of one window (shared memory: (kSelOut + window + 8) keys)

thread_local char* g_err = nullptr;
thread_local size_t g_err_cap = 0;
int sel_fail(int code, const char* fmt, ...) {
  if (!g_err || !g_err_cap) return code;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, g_err_cap, fmt, ap);
  va_end(ap);
  return code;
}

// The answer is found as an order-preserving 32-bit KEY (larger float <=> larger unsigned key: sign bit flipped for
// non-negative floats, all bits for negative ones); val_of maps a key back to the float with that rank.
__device__ __forceinline__ float val_of(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// one 128-bit group of samples against the four candidates: element e of the group sits at offset d0 + e from the thread's first
// output; output q counts offsets q <= d <= q + W - 1 (EDGE: test it; interior groups satisfy it for every q).  The compare is
// a float compare and the counter a float (exact up to 2^24): FSET on the ALU pipe + FADD on the FMA pipe, where an integer
// compare-and-count is two ALU-pipe instructions -- the kernel is bound by that pipe.
template <bool EDGE>
__device__ __forceinline__ void count_group(const float4 kv, int d0, int W, const float (&thr)[4], float (&cnt)[4]) {
  const float ke[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      bool in = ke[e] >= thr[q];
      if (EDGE) in = in && (d0 + e >= q) && (d0 + e <= q + W - 1);
      cnt[q] += in ? 1.f : 0.f;
    }
  }
}

// float whose order among floats is the key's order among keys; keys below key(-inf) (negative-NaN patterns) count everything,
// like -inf; keys above key(+inf) are NaN and count nothing (no sample is above +inf: NaN samples are handled apart)
__device__ __forceinline__ float thr_of(unsigned key) { return val_of(key < 0x007fffffu ? 0x007fffffu : key); }

// grid: one CTA per (trace, tile of kSelOut outputs), linearised
SEL_KERNEL(kSelThreads) kth_window_fwd(const float* __restrict__ y, float* __restrict__ out, int* __restrict__ arg,
                                       int T, int a, int W, int k, int tiles_per_trace) {
  SEL_DYN_SMEM(float, vals);
  SEL_SHARED(int, has_nan);
  const long long trace = blockIdx.x / tiles_per_trace;
  const int tb = (blockIdx.x - trace * tiles_per_trace) * kSelOut;  // first output of the tile
  const float* yt = y + trace * T;
  if (threadIdx.x == 0) has_nan = 0;
  __syncthreads();
  // stage samples tb + a .. tb + a + kSelOut + W + 2 (-0 read as +0; NaN past the horizon: no compare counts it)
  const int n_stage = (kSelOut + W + 3 + 3) & ~3;
  for (int i = threadIdx.x; i < n_stage; i += kSelThreads) {
    const long long s = (long long)tb + a + i;
    float v = __int_as_float(0x7fc00000);
    if (s < T) {
      v = yt[s] + 0.f;  // (-0) + (+0) = +0
      if (v != v) has_nan = 1;  // (an operand value the reference raises on)
    }
    vals[i] = v;
  }
  __syncthreads();
  const int l0 = 4 * threadIdx.x;  // local index of the thread's first output = of the first sample of its first window
  const float4* k4 = reinterpret_cast<const float4*>(vals) + threadIdx.x;
  const int g_last = (W + 2) >> 2;                 // last group that holds a sample of the 4 windows (offsets 0 .. W + 2)
  const int g_int = W >= 4 ? ((W - 4) >> 2) : 0;   // groups 1 .. g_int lie inside every one of the 4 windows
  const float kf = (float)k;
  unsigned pre[4] = {0u, 0u, 0u, 0u};
  for (int bit = 31; bit >= 0; --bit) {
    float thr[4];
    float cnt[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) thr[q] = thr_of(pre[q] | (1u << bit));
    count_group<true>(k4[0], 0, W, thr, cnt);
    int g = 1;
#pragma unroll 2
    for (; g <= g_int; ++g) count_group<false>(k4[g], 4 * g, W, thr, cnt);
    for (; g <= g_last; ++g) count_group<true>(k4[g], 4 * g, W, thr, cnt);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (cnt[q] >= kf) pre[q] |= 1u << bit;
  }
  float ans[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) ans[q] = val_of(pre[q]);
  // occurrence number of the answer among its equals (stable ascending sort, position n - k): #(sample >= answer) - k
  int occ[4] = {0, 0, 0, 0};
  if (arg) {
    float cge[4] = {0.f, 0.f, 0.f, 0.f};
    for (int g = 0; g <= g_last; ++g) count_group<true>(k4[g], 4 * g, W, ans, cge);
#pragma unroll
    for (int q = 0; q < 4; ++q) occ[q] = (int)cge[q] - k;
  }
  const bool any_nan = has_nan != 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int t = tb + l0 + q;
    if (t >= T) break;
    long long last = (long long)t + a + W - 1;
    if (last > T - 1) last = T - 1;
    const int n = (int)(last - ((long long)t + a) + 1);
    bool nan_in = false;
    if (any_nan)  // rare: a NaN inside the window (a value the reference raises on, api.py:218-221) makes the output NaN
      for (int d = 0; d < n; ++d) nan_in = nan_in || vals[l0 + q + d] != vals[l0 + q + d];
    const bool has = n >= k && !nan_in;
    out[trace * T + t] = nan_in ? __int_as_float(0x7fc00000) : has ? ans[q] : __int_as_float(0xff800000);
    if (arg) {
      int where = -1;
      if (has) {
        int seen = 0;
        for (int d = 0; d < n; ++d) {
          if (vals[l0 + q + d] == ans[q]) {
            if (seen == occ[q]) {
              where = t + a + d;
              break;
            }
            ++seen;
          }
        }
      }
      arg[trace * T + t] = where;
    }
  }
}

// grad_y[b, s] = sum of grad_out[b, t] over the outputs t in [s - a - W + 1, s - a] that selected s
SEL_KERNEL(256) kth_window_bwd(const int* __restrict__ arg, const float* __restrict__ grad_out,
                               float* __restrict__ grad_y, long long B, int T, int a, int W) {
  const long long n = B * (long long)T;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n; e += (long long)gridDim.x * 256) {
    const long long trace = e / T;
    const int s = (int)(e - trace * T);
    long long t1 = (long long)s - a, t0 = t1 - W + 1;
    if (t0 < 0) t0 = 0;
    float g = 0.f;
    const int* ar = arg + trace * T;
    const float* go = grad_out + trace * T;
    for (long long t = t0; t <= t1; ++t)
      if (ar[t] == s) g += go[t];
    grad_y[e] = g;
  }
}

// Until with the exclusive prefix: one thread per output walks its window (consecutive threads read consecutive samples).
// key: 2 * sample + 1 when the value is an r, 2 * sample when it is an l, -1 where there is no value.
SEL_KERNEL(256) ctstl_until_fwd(const float* __restrict__ l, const float* __restrict__ r, float* __restrict__ out,
                                int* __restrict__ key, long long B, int T, int a, int b_end) {
  const long long n = B * (long long)T;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n; e += (long long)gridDim.x * 256) {
    const long long trace = e / T;
    const int t = (int)(e - trace * T);
    const float* lt = l + trace * T + t;
    const float* rt = r + trace * T + t;
    int last = T - 1 - t;
    if (b_end >= 0 && b_end < last) last = b_end;
    if (a > last) {  // empty window: the reference raises (ctstl.py:96-97)
      out[e] = __int_as_float(0x7fc00000);
      if (key) key[e] = -1;
      continue;
    }
    float P = __int_as_float(0x7f800000);  // min of l[t .. t + i - 1]
    int argP = -1;
    for (int i = 0; i < a; ++i) {
      const float v = lt[i];
      if (v < P) { P = v; argP = i; }
    }
    float best = __int_as_float(0xff800000);
    int bk = -1;
    for (int i = a; i <= last; ++i) {
      const float rv = rt[i];
      const bool from_r = rv <= P;  // min(r, prefix): ties go to r
      const float cand = from_r ? rv : P;
      if (cand > best || bk < 0) {
        best = cand;
        bk = from_r ? 2 * (t + i) + 1 : 2 * (t + argP);
      }
      const float v = lt[i];
      if (v < P) { P = v; argP = i; }
    }
    out[e] = best;
    if (key) key[e] = bk;
  }
}

// grad_l[s] collects from the outputs t in [s - b_end + 1, s] whose key is 2 s; grad_r[s] from t in [s - b_end, s - a], key 2 s + 1
SEL_KERNEL(256) ctstl_until_bwd(const int* __restrict__ key, const float* __restrict__ grad_out,
                                float* __restrict__ grad_l, float* __restrict__ grad_r, long long B, int T, int a, int b_end) {
  const long long n = B * (long long)T;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n; e += (long long)gridDim.x * 256) {
    const long long trace = e / T;
    const int s = (int)(e - trace * T);
    const int* kt = key + trace * T;
    const float* go = grad_out + trace * T;
    int t0 = b_end >= 0 ? s - b_end : 0;
    if (t0 < 0) t0 = 0;
    float gl = 0.f, gr = 0.f;
    for (int t = t0; t <= s; ++t) {
      const int kk = kt[t];
      if (kk == 2 * s) gl += go[t];
      if (kk == 2 * s + 1) gr += go[t];
    }
    grad_l[e] = gl;
    grad_r[e] = gr;
  }
}

#if !defined(PYSTL_SELECT_EMU)
int check_shape(long long B, long long T) {
  if (B < 0 || T < 0 || T >= (1ll << 30)) return sel_fail(PYSTL_E_INVALID, "bad shape");
  return PYSTL_OK;
}
int need_device() {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return sel_fail(PYSTL_E_CUDA, "no CUDA device: the b200 backend has no CPU path");
  }
  return PYSTL_OK;
}
unsigned flat_grid(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = 148ll * 64;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
// window length of [a, b_end] on a trace of T samples (b_end < 0: open, clipped at the horizon)
long long window_of(int a, int b_end, long long T) {
  long long W = b_end >= 0 ? (long long)b_end - a + 1 : T - a;
  if (W > T) W = T;
  return W;
}

#endif  // !PYSTL_SELECT_EMU
}  // namespace

#if !defined(PYSTL_SELECT_EMU)
int ctstl_cumulative_forward(const float* y, int64_t B, int64_t T, int32_t a, int32_t b_end, int32_t k, float* out, int32_t* arg,
                             void* stream, char* err, size_t err_cap) {
  g_err = err;
  g_err_cap = err_cap;
  int rc = check_shape(B, T);
  if (rc) return rc;
  if (a < 0) return sel_fail(PYSTL_E_INVALID, "Interval start must be >= 0.");
  if (b_end >= 0 && b_end < a) return sel_fail(PYSTL_E_INVALID, "Interval end must be >= start.");
  if (k < 1) return sel_fail(PYSTL_E_INVALID, "tau must be > 0.");  // k = ceil(tau / delta), ctstl.py:28-35
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!y || !out) return sel_fail(PYSTL_E_INVALID, "null buffer");
  long long W = window_of(a, b_end, T);
  if (W < 1) W = 1;  // every window is empty: -inf everywhere, one pass over nothing
  if (W > kSelMaxWindow) return sel_fail(PYSTL_E_UNSUPPORTED, "C^tau_I over a window of %lld samples (at most %d)", W, kSelMaxWindow);
  if ((rc = need_device())) return rc;
  const long long tiles = (T + kSelOut - 1) / kSelOut;
  if (B * tiles >= (1ll << 31)) return sel_fail(PYSTL_E_UNSUPPORTED, "too many tiles for one launch");
  const size_t smem = (size_t)((kSelOut + W + 3 + 3) & ~3ll) * sizeof(unsigned);
  static thread_local size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    const cudaError_t e = cudaFuncSetAttribute(kth_window_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return sel_fail(PYSTL_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    smem_set = 227 * 1024;
  }
  kth_window_fwd<<<(unsigned)(B * tiles), kSelThreads, smem, (cudaStream_t)stream>>>(y, out, arg, (int)T, (int)a, (int)W, (int)k, (int)tiles);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return sel_fail(PYSTL_E_CUDA, "kth_window_fwd: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

int ctstl_cumulative_backward(const int32_t* arg, const float* grad_out, int64_t B, int64_t T, int32_t a, int32_t b_end, float* grad_y,
                              void* stream, char* err, size_t err_cap) {
  g_err = err;
  g_err_cap = err_cap;
  int rc = check_shape(B, T);
  if (rc) return rc;
  if (a < 0) return sel_fail(PYSTL_E_INVALID, "Interval start must be >= 0.");
  if (b_end >= 0 && b_end < a) return sel_fail(PYSTL_E_INVALID, "Interval end must be >= start.");
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!arg || !grad_out || !grad_y) return sel_fail(PYSTL_E_INVALID, "null buffer");
  if ((rc = need_device())) return rc;
  long long W = window_of(a, b_end, T);
  if (W < 1) W = 1;
  kth_window_bwd<<<flat_grid(B * T), 256, 0, (cudaStream_t)stream>>>(arg, grad_out, grad_y, (long long)B, (int)T, (int)a, (int)W);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return sel_fail(PYSTL_E_CUDA, "kth_window_bwd: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

int ctstl_until_forward(const float* l, const float* r, int64_t B, int64_t T, int32_t a, int32_t b_end, float* out, int32_t* key,
                        void* stream, char* err, size_t err_cap) {
  g_err = err;
  g_err_cap = err_cap;
  int rc = check_shape(B, T);
  if (rc) return rc;
  if (a < 0) return sel_fail(PYSTL_E_INVALID, "UNTIL start must be >= 0.");  // ctstl.py:92-93
  if (b_end >= 0 && b_end < a) return sel_fail(PYSTL_E_INVALID, "Interval end must be >= start.");
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!l || !r || !out) return sel_fail(PYSTL_E_INVALID, "null buffer");
  if ((rc = need_device())) return rc;
  ctstl_until_fwd<<<flat_grid(B * T), 256, 0, (cudaStream_t)stream>>>(l, r, out, key, (long long)B, (int)T, (int)a, (int)b_end);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return sel_fail(PYSTL_E_CUDA, "ctstl_until_fwd: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

int ctstl_until_backward(const int32_t* key, const float* grad_out, int64_t B, int64_t T, int32_t a, int32_t b_end, float* grad_l,
                         float* grad_r, void* stream, char* err, size_t err_cap) {
  g_err = err;
  g_err_cap = err_cap;
  int rc = check_shape(B, T);
  if (rc) return rc;
  if (a < 0) return sel_fail(PYSTL_E_INVALID, "UNTIL start must be >= 0.");
  if (b_end >= 0 && b_end < a) return sel_fail(PYSTL_E_INVALID, "Interval end must be >= start.");
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!key || !grad_out || !grad_l || !grad_r) return sel_fail(PYSTL_E_INVALID, "null buffer");
  if ((rc = need_device())) return rc;
  ctstl_until_bwd<<<flat_grid(B * T), 256, 0, (cudaStream_t)stream>>>(key, grad_out, grad_l, grad_r, (long long)B, (int)T, (int)a, (int)b_end);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return sel_fail(PYSTL_E_CUDA, "ctstl_until_bwd: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

#endif  // !PYSTL_SELECT_EMU

}  // namespace pystl
