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
#if defined(PYSTL_SCAN_EMU)
// Host build for the CPU test tier (tests/emu/scan_emu.cpp): the kernels below are compiled with g++, every CUDA thread a fiber of
// tests/emu/simt_host.h (warp collectives are cooperative yields).  Test infrastructure; the product is the CUDA build.
#include "scan_emu_shim.h"
#else
#include <cuda_runtime.h>
#endif

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
  static __device__ A push(A l, A acc) { return R::comb(l, acc); }
  static __device__ float fin(A a, float n, const RP& rp) { return R::fin(a, n, rp).x[0]; }
  static __device__ int arg(A) { return 0; }
};
template <int MODE>
struct ScanRed<SEM_CUMULATIVE, MODE> {  // one channel (pos or neg) of a cumulative value: Eventually SUMS the window (cumulative.py:62-73)
  using A = Vec<1>;                     // (Always is the minimum per channel, cumulative.py:35-46: the Classic kernels serve it)
  static __device__ A ident() { return A{{0.f}}; }
  static __device__ A lift(float y, int, const RP&) { return A{{y}}; }
  static __device__ A comb(A a, A b) { return A{{a.x[0] + b.x[0]}}; }
  static __device__ A push(A l, A acc) { return A{{l.x[0] + acc.x[0]}}; }
  static __device__ float fin(A a, float, const RP&) { return a.x[0]; }
  static __device__ int arg(A) { return 0; }
};
template <int MODE>
struct ScanRed<SEM_DGMSR, MODE> {  // the formula kernels' reducer; fin BRANCHES on the flag instead of evaluating both means
  using R = Red<SEM_DGMSR, MODE>;  // (a long suffix almost always holds a non-positive entry: rcp + sqrt instead of seven MUFU ops)
  using A = typename R::A;
  static __device__ A ident() { return R::ident(); }
  static __device__ A lift(float y, int, const RP& rp) { return R::lift(Vec<1>{{y}}, rp, 1.f); }
  static __device__ A comb(A a, A b) { return R::comb(a, b); }
  static __device__ A push(A l, A acc) { return R::comb(l, acc); }
  static __device__ float fin(A a, float n, const RP& rp) {
    float h;
    if (a.x[1] == f_inf()) {  // _dgmsr.py:83-87
      const float m = rp.eps_p + a.x[0] * fast_rcp(n);
      h = rp.sqrt_eps - (PYSTL_P(rp) == 1 ? fast_sqrt(m) : fast_pow(m, rp.inv2p));
    } else {
      h = R::fin(a, n, rp).x[0];
      return h;
    }
    return MODE == MODE_AND ? h : -h;
  }
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
  static __device__ A push(A l, A acc) { return comb(l, acc); }
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
  // one element in front of an accumulator: one MUFU instead of two
  static __device__ A push(A l, A acc) {
    const float d = l.x[0] - acc.x[0];
    const float ex = fast_ex2(-fabsf(d));
    return A{{fmaxf(l.x[0], acc.x[0]), d <= 0.f ? fmaf(l.x[1], ex, acc.x[1]) : fmaf(acc.x[1], ex, l.x[1])}};
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
template <class A>
__device__ A shfl_acc(A v, int src) {
  A r;
#pragma unroll
  for (int w = 0; w < (int)(sizeof(A) / sizeof(float)); ++w) r.x[w] = __shfl_sync(0xffffffffu, v.x[w], src);
  return r;
}

// A warp walks its trace in CHUNKS of 256 samples: global loads and stores are coalesced rows of 32 consecutive samples,
// staged through a per-warp shared-memory buffer so that every lane then owns 8 CONSECUTIVE samples (one pad word per 32
// keeps both access patterns conflict-free).  The trace is read once and written once.
constexpr int kChunk = 256;
constexpr int kBuf = kChunk + kChunk / 32;
__device__ __forceinline__ int pad(int e) { return e + (e >> 5); }

template <int SEM, int MODE>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) open_window_fwd(const float* __restrict__ y, float* __restrict__ out,
                                                                       int* __restrict__ arg, long long B, int T, int a, RP rp,
                                                                       int nan_tail) {
  using S = ScanRed<SEM, MODE>;
  using A = typename S::A;
  __shared__ float s_val[kWarpsPerBlock][kBuf];
  __shared__ int s_arg[SEM == SEM_CLASSIC ? kWarpsPerBlock : 1][SEM == SEM_CLASSIC ? kBuf : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * kWarpsPerBlock + warp;
  if (b >= B) return;
  const float* yb = y + b * T;
  float* ob = out + b * T;
  int* ab = arg ? arg + b * T : nullptr;
  float* buf = s_val[warp];
  A carry = S::ident();  // everything after the current chunk
  const int c_last = ((T - 1) / kChunk) * kChunk;
  float nxt[8];  // the next chunk's rows, in flight while this one is reduced
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int u = c_last + j * 32 + lane;
    nxt[j] = u < T ? yb[u] : 0.f;
  }
  for (int c0 = c_last; c0 >= 0; c0 -= kChunk) {
#pragma unroll
    for (int j = 0; j < 8; ++j) buf[pad(j * 32 + lane)] = nxt[j];
    if (c0 > 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) nxt[j] = yb[c0 - kChunk + j * 32 + lane];
    }
    __syncwarp();
    const int u0 = c0 + lane * 8;
    A l8[8];
    A tot = S::ident();
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      l8[i] = S::lift(buf[pad(lane * 8 + i)], u0 + i, rp);
      if (u0 + i < T) tot = S::push(l8[i], tot);
    }
    A inc = tot;  // suffix over the lanes: this lane's samples and every later lane's
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const A other = shfl_down_acc(inc, d);
      if (lane + d < 32) inc = S::comb(inc, other);
    }
    A run = shfl_down_acc(inc, 1);
    run = lane == 31 ? carry : S::comb(run, carry);
    carry = S::comb(shfl_acc(inc, 0), carry);
    float last = 0.f;  // the value of the window [T - 1, T - 1]
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      const int u = u0 + i;
      if (u < T) {
        run = S::push(l8[i], run);
        const float v = S::fin(run, (float)(T - u), rp);
        buf[pad(lane * 8 + i)] = v;
        if constexpr (SEM == SEM_CLASSIC) s_arg[warp][pad(lane * 8 + i)] = S::arg(run);
        if (u == T - 1) last = v;
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = j * 32 + lane, u = c0 + e, t = u - a;
      if (u < T && t >= 0) {
        ob[t] = buf[pad(e)];
        if constexpr (SEM == SEM_CLASSIC) ab[t] = s_arg[warp][pad(e)];
      }
    }
    if (c0 + kChunk >= T && a > 0) {
      // t + a > T - 1: the reference raises (masked by the caller's start_slack); a finite value of the right magnitude for a
      // parent kernel to read, NaN (mode bit 1) when these values ARE the formula's robustness
      const float v = nan_tail ? __int_as_float(0x7fc00000) : __shfl_sync(0xffffffffu, last, ((T - 1 - c0) >> 3));
      for (int tt = (T - a > 0 ? T - a : 0) + lane; tt < T; tt += 32) {
        ob[tt] = v;
        if constexpr (SEM == SEM_CLASSIC) ab[tt] = -1;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- classical Until without an end
// W(u) = min(l_u, max(r_u, W(u + 1))), W(T) = -inf (classic.py:59-80 with Interval.end None): the composition of the clamps
// f_u(w) = min(l_u, max(r_u, w)) is a clamp again (tile_eval.h: ClampOp, the operator of the in-kernel clamp scan), so the
// unbounded Until is the same chunked suffix scan over (A, src A, B, src B) quadruples.  out[t] = W(t + a); the host adds
// G[0, a] l around it (lower.cpp: U[a, inf) = min(G[0,a] l, W(t + a))).  key[t] = 2 * sample + (1 if the value is an r).
__global__ void __launch_bounds__(32 * kWarpsPerBlock) open_until_fwd_classic(const float* __restrict__ l, const float* __restrict__ r,
                                                                              float* __restrict__ out, int* __restrict__ key, long long B,
                                                                              int T, int a) {
  using OP = ClampOp<SEM_CLASSIC_SRC>;
  using A = OP::A;
  __shared__ float s_l[kWarpsPerBlock][kBuf], s_r[kWarpsPerBlock][kBuf];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * kWarpsPerBlock + warp;
  if (b >= B) return;
  const float* lb = l + b * T;
  const float* rb = r + b * T;
  float* ob = out + b * T;
  int* kb = key + b * T;
  float *bl = s_l[warp], *br = s_r[warp];
  A carry = OP::ident();
  for (int c0 = ((T - 1) / kChunk) * kChunk; c0 >= 0; c0 -= kChunk) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = j * 32 + lane, u = c0 + e;
      bl[pad(e)] = u < T ? lb[u] : 0.f;
      br[pad(e)] = u < T ? rb[u] : 0.f;
    }
    __syncwarp();
    const int u0 = c0 + lane * 8;
    A x8[8];
    A tot = OP::ident();
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      const int u = u0 + i;
      x8[i] = u < T ? OP::make(Vec<2>{{bl[pad(lane * 8 + i)], i2f(2 * u)}}, Vec<2>{{br[pad(lane * 8 + i)], i2f(2 * u + 1)}}) : OP::ident();
      tot = OP::comb(x8[i], tot);
    }
    A inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const A other = shfl_down_acc(inc, d);
      if (lane + d < 32) inc = OP::comb(inc, other);
    }
    A run = shfl_down_acc(inc, 1);
    run = lane == 31 ? carry : OP::comb(run, carry);
    carry = OP::comb(shfl_acc(inc, 0), carry);
    float last = 0.f;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      const int u = u0 + i;
      if (u < T) {
        run = OP::comb(x8[i], run);
        const Vec<2> v = OP::value(run);
        bl[pad(lane * 8 + i)] = v.x[0];
        br[pad(lane * 8 + i)] = v.x[1];
        if (u == T - 1) last = v.x[0];
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = j * 32 + lane, u = c0 + e, t = u - a;
      if (u < T && t >= 0) {
        ob[t] = bl[pad(e)];
        kb[t] = f2i(br[pad(e)]);
      }
    }
    if (c0 + kChunk >= T && a > 0) {
      const float v = __shfl_sync(0xffffffffu, last, ((T - 1 - c0) >> 3));
      for (int tt = (T - a > 0 ? T - a : 0) + lane; tt < T; tt += 32) {
        ob[tt] = v;
        kb[tt] = -1;
      }
    }
    __syncwarp();
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
  static __device__ A push(A acc, A l) { return comb(acc, l); }
  static __device__ float apply(A w, float y, const RP& rp) { return AD::in_factor(w, y, 1.f, rp); }
};
template <int MODE>
struct ScanAdj<SEM_CUMULATIVE, MODE> {  // d sum / d y_k = 1: the prefix sum of the cotangents
  using A = Vec<1>;
  static __device__ A ident() { return A{{0.f}}; }
  static __device__ A coeff(float g, float, float, const RP&) { return A{{g}}; }
  static __device__ A comb(A a, A b) { return A{{a.x[0] + b.x[0]}}; }
  static __device__ A push(A acc, A l) { return A{{acc.x[0] + l.x[0]}}; }
  static __device__ float apply(A w, float, const RP&) { return w.x[0]; }
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
  static __device__ A push(A acc, A l) {  // one coefficient behind an accumulator: one MUFU
    if (l.x[0] == -f_inf()) return acc;
    const float d = l.x[0] - acc.x[0];
    const float ex = fast_ex2(-fabsf(d));
    return A{{fmaxf(l.x[0], acc.x[0]), d <= 0.f ? fmaf(l.x[1], ex, acc.x[1]) : fmaf(acc.x[1], ex, l.x[1])}};
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
  __shared__ float s_g[kWarpsPerBlock][kBuf], s_o[kWarpsPerBlock][kBuf], s_y[kWarpsPerBlock][kBuf];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * kWarpsPerBlock + warp;
  if (b >= B) return;
  const float* yb = y + b * T;
  const float* ob = out + b * T;
  const float* gb = grad_out + b * T;
  float* dy = grad_y + b * T;
  float *bg = s_g[warp], *bo = s_o[warp], *by = s_y[warp];
  const int nvalid = T - a;  // outputs t < nvalid exist; output t serves the samples k >= t + a
  for (int k = lane; k < a && k < T; k += 32) dy[k] = 0.f;  // samples before the first window
  A carry = S::ident();  // every output before the current chunk
  for (int c0 = 0; c0 < nvalid; c0 += kChunk) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = j * 32 + lane, t = c0 + e;
      const bool live = t < nvalid;
      bg[pad(e)] = live ? gb[t] : 0.f;
      bo[pad(e)] = live ? ob[t] : 0.f;
      by[pad(e)] = live ? yb[t + a] : 0.f;
    }
    __syncwarp();
    const int t0 = c0 + lane * 8;
    A k8[8];
    A tot = S::ident();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = t0 + i;
      k8[i] = t < nvalid ? S::coeff(bg[pad(lane * 8 + i)], bo[pad(lane * 8 + i)], (float)(T - t - a), rp) : S::ident();
      tot = S::push(tot, k8[i]);
    }
    A inc = tot;  // prefix over the lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const A other = shfl_up_acc(inc, d);
      if (lane >= d) inc = S::comb(other, inc);
    }
    A run = shfl_up_acc(inc, 1);
    run = lane == 0 ? carry : S::comb(carry, run);
    carry = S::comb(carry, shfl_acc(inc, 31));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      run = S::push(run, k8[i]);
      bg[pad(lane * 8 + i)] = S::apply(run, by[pad(lane * 8 + i)], rp);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = j * 32 + lane, t = c0 + e;
      if (t < nvalid) dy[t + a] = bg[pad(e)];
    }
    __syncwarp();
  }
}
// classic: every output routes its cotangent to the recorded arg-min / arg-max.  The arg of a suffix window is
// non-decreasing in t, so the outputs that share a target are one RUN of consecutive t: a segmented prefix sum of the
// cotangents, written once at the end of each run (no atomics; grad_y is zeroed first).
struct Seg {
  int f;    // a run starts inside
  float v;  // sum since the last start (the whole range when none)
};
__device__ __forceinline__ Seg seg_comb(Seg a, Seg b) { return Seg{a.f | b.f, b.f ? b.v : a.v + b.v}; }

template <bool TWO>  // TWO: the targets are keys 2 * sample + side of an Until (side 0: grad_y, 1: grad_r)
__global__ void __launch_bounds__(32 * kWarpsPerBlock) open_window_bwd_classic(const int* __restrict__ arg, const float* __restrict__ grad_out,
                                                                               float* __restrict__ grad_y, float* __restrict__ grad_r,
                                                                               long long B, int T, int a) {
  __shared__ float s_g[kWarpsPerBlock][kBuf];
  __shared__ int s_a[kWarpsPerBlock][kBuf + 2];
  // TWO: the runs of an Until are short (the source is often the sample itself), one 4-byte store per sample; targets inside
  // [c0 + a, c0 + a + 256) are staged here and leave as coalesced rows (zeros are not stored: grad is zero-filled already)
  __shared__ float s_t[TWO ? 2 : 1][TWO ? kWarpsPerBlock : 1][TWO ? kBuf : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * kWarpsPerBlock + warp;
  if (b >= B) return;
  const int* ab = arg + b * T;
  const float* gb = grad_out + b * T;
  float* dy = grad_y + b * T;
  float* dr = TWO ? grad_r + b * T : nullptr;
  float* bg = s_g[warp];
  int* ba = s_a[warp];
  const int nvalid = T - a;
  float carry = 0.f;  // the open run's sum over the chunks before this one
  int prev = -1;      // arg of the sample before the chunk
  for (int c0 = 0; c0 < nvalid; c0 += kChunk) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = j * 32 + lane, t = c0 + e;
      bg[pad(e)] = t < nvalid ? gb[t] : 0.f;
      ba[pad(e)] = t < nvalid ? ab[t] : -2;
      if constexpr (TWO) s_t[0][warp][pad(e)] = s_t[1][warp][pad(e)] = 0.f;
    }
    const int after = c0 + kChunk < nvalid ? ab[c0 + kChunk] : -2;  // (uniform load)
    __syncwarp();
    int k[10];  // this lane's 8 targets and the neighbours on either side
#pragma unroll
    for (int i = 0; i < 8; ++i) k[i + 1] = ba[pad(lane * 8 + i)];
    k[0] = lane == 0 ? prev : ba[pad(lane * 8 - 1)];
    k[9] = lane == 31 ? after : ba[pad(lane * 8 + 8)];
    float g8[8];
    Seg tot{0, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      g8[i] = bg[pad(lane * 8 + i)];
      tot = seg_comb(tot, Seg{k[i + 1] != k[i], g8[i]});
    }
    Seg inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      Seg other{__shfl_up_sync(0xffffffffu, inc.f, d), __shfl_up_sync(0xffffffffu, inc.v, d)};
      if (lane >= d) inc = seg_comb(other, inc);
    }
    Seg ex{__shfl_up_sync(0xffffffffu, inc.f, 1), __shfl_up_sync(0xffffffffu, inc.v, 1)};
    if (lane == 0) ex = Seg{0, 0.f};
    float run = ex.f ? ex.v : carry + ex.v;
    const Seg whole{__shfl_sync(0xffffffffu, inc.f, 31), __shfl_sync(0xffffffffu, inc.v, 31)};
    carry = whole.f ? whole.v : carry + whole.v;
    prev = __shfl_sync(0xffffffffu, k[8], 31);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (k[i + 1] != k[i]) run = 0.f;
      run += g8[i];
      if (k[i + 1] >= 0 && k[i + 2] != k[i + 1]) {  // the run ends here
        if constexpr (TWO) {
          const int kk = k[i + 1] >> 1, side = k[i + 1] & 1, rel = kk - (c0 + a);
          if (rel < kChunk) s_t[side][warp][pad(rel)] = run;  // (rel >= 0: a source is never before its output's t + a)
          else (side ? dr : dy)[kk] = run;
        } else {
          dy[k[i + 1]] = run;
        }
      }
    }
    __syncwarp();
    if constexpr (TWO) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = j * 32 + lane;
        const float v0 = s_t[0][warp][pad(e)], v1 = s_t[1][warp][pad(e)];
        if (v0 != 0.f) dy[c0 + a + e] = v0;
        if (v1 != 0.f) dr[c0 + a + e] = v1;
      }
      __syncwarp();
    }
  }
}

#if !defined(PYSTL_SCAN_EMU)
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

#endif  // !PYSTL_SCAN_EMU

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

#if !defined(PYSTL_SCAN_EMU)
int check(int semantics, int mode, int a, long long B, long long T, float temperature, float eps, int p) {
  if (semantics != SEM_CLASSIC && semantics != SEM_SMOOTH && semantics != SEM_AGM && semantics != SEM_DGMSR && semantics != SEM_CUMULATIVE)
    return scan_fail(PYSTL_E_INVALID, "unknown semantics %d", semantics);
  if (mode < 0 || mode > 3) return scan_fail(PYSTL_E_INVALID, "mode must be 0 (always) or 1 (eventually), + 2 for a NaN tail");
  if (a < 0) return scan_fail(PYSTL_E_INVALID, "Interval start must be >= 0.");
  if (B < 0 || T < 0 || T >= (1ll << 31) - 65536) return scan_fail(PYSTL_E_INVALID, "bad shape");
  if (semantics == SEM_SMOOTH && !(temperature > 0.f)) return scan_fail(PYSTL_E_INVALID, "temperature must be > 0.");
  if (semantics == SEM_DGMSR && (!(eps > 0.f) || p < 1)) return scan_fail(PYSTL_E_INVALID, "dgmsr needs eps > 0 and p >= 1");
  return PYSTL_OK;
}
#endif  // !PYSTL_SCAN_EMU

}  // namespace

#if !defined(PYSTL_SCAN_EMU)
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
  if (semantics == SEM_CUMULATIVE && (mode & 1) == 0) semantics = SEM_CLASSIC;  // Always: the minimum of the channel
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!y || !out) return scan_fail(PYSTL_E_INVALID, "null buffer");
  if (semantics == SEM_CLASSIC && !arg) return scan_fail(PYSTL_E_INVALID, "a minimum / maximum needs the arg buffer");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return scan_fail(PYSTL_E_CUDA, "no CUDA device: the b200 backend has no CPU path");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const RP rp = make_rp(temperature, eps, p);
  int* argp = semantics == SEM_CLASSIC ? arg : nullptr;
  const int nan_tail = (mode >> 1) & 1;
  mode &= 1;
  if (semantics == SEM_CUMULATIVE) {  // (y is one channel of the pairs) Eventually: the window's sum
    const unsigned grid = (unsigned)((B + kWarpsPerBlock - 1) / kWarpsPerBlock);
    open_window_fwd<SEM_CUMULATIVE, MODE_OR><<<grid, 32 * kWarpsPerBlock, 0, st>>>(y, out, nullptr, (long long)B, (int)T, (int)a, rp, nan_tail);
  } else {
    PYSTL_SCAN_DISPATCH(open_window_fwd, y, out, argp, (long long)B, (int)T, (int)a, rp, nan_tail);
  }
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
  mode &= 1;
  if (semantics == SEM_CUMULATIVE && mode == 0) semantics = SEM_CLASSIC;  // Always: the minimum of the channel
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
    open_window_bwd_classic<false><<<(unsigned)((B + kWarpsPerBlock - 1) / kWarpsPerBlock), 32 * kWarpsPerBlock, 0, st>>>(
        arg, grad_out, grad_y, nullptr, (long long)B, (int)T, (int)a);
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
    else if (semantics == SEM_CUMULATIVE)
      open_window_bwd<SEM_CUMULATIVE, MODE_OR><<<grid, block, 0, st>>>(y, out, grad_out, grad_y, (long long)B, (int)T, (int)a, rp);
    else PYSTL_BWD(SEM_DGMSR);
#undef PYSTL_BWD
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return scan_fail(PYSTL_E_CUDA, "open_window_bwd: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

int open_until_forward(int32_t semantics, int32_t a, const float* l, const float* r, int64_t B, int64_t T, float* out, int32_t* key,
                       void* stream, char* err, size_t err_cap) {
  g_err = err;
  g_err_cap = err_cap;
  if (semantics != SEM_CLASSIC)
    return scan_fail(PYSTL_E_UNSUPPORTED, "an Until without an interval end beyond one tile: classical semantics only (the others are O(T^2))");
  if (a < 0) return scan_fail(PYSTL_E_INVALID, "Interval start must be >= 0.");
  if (B < 0 || T < 0 || T >= (1ll << 30)) return scan_fail(PYSTL_E_INVALID, "bad shape");
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!l || !r || !out || !key) return scan_fail(PYSTL_E_INVALID, "null buffer");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return scan_fail(PYSTL_E_CUDA, "no CUDA device: the b200 backend has no CPU path");
  }
  open_until_fwd_classic<<<(unsigned)((B + kWarpsPerBlock - 1) / kWarpsPerBlock), 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>(
      l, r, out, key, (long long)B, (int)T, (int)a);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return scan_fail(PYSTL_E_CUDA, "open_until_fwd: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

int open_until_backward(int32_t semantics, int32_t a, const int32_t* key, const float* grad_out, int64_t B, int64_t T, float* grad_l,
                        float* grad_r, void* stream, char* err, size_t err_cap) {
  g_err = err;
  g_err_cap = err_cap;
  if (semantics != SEM_CLASSIC)
    return scan_fail(PYSTL_E_UNSUPPORTED, "an Until without an interval end beyond one tile: classical semantics only (the others are O(T^2))");
  if (a < 0) return scan_fail(PYSTL_E_INVALID, "Interval start must be >= 0.");
  if (B < 0 || T < 0 || T >= (1ll << 30)) return scan_fail(PYSTL_E_INVALID, "bad shape");
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!key || !grad_out || !grad_l || !grad_r) return scan_fail(PYSTL_E_INVALID, "null buffer");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return scan_fail(PYSTL_E_CUDA, "no CUDA device: the b200 backend has no CPU path");
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(grad_l, 0, (size_t)B * (size_t)T * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(grad_r, 0, (size_t)B * (size_t)T * sizeof(float), st);
  if (e != cudaSuccess) return scan_fail(PYSTL_E_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  open_window_bwd_classic<true><<<(unsigned)((B + kWarpsPerBlock - 1) / kWarpsPerBlock), 32 * kWarpsPerBlock, 0, st>>>(
      key, grad_out, grad_l, grad_r, (long long)B, (int)T, (int)a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return scan_fail(PYSTL_E_CUDA, "open_until_bwd: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

// ---------------------------------------------------------------- the tensor a formula kernel reads next to leaf channels
// out[r, :] = [ signal[r, ch[0..n)] | col_0[r] .. col_{L-1}[r] ] for every row r = (trace, t): one pass (a torch.cat of an
// indexed signal and stacked columns is three), and the adjoint: grad_signal[r, :] = 0 except the read channels, one
// gradient column per leaf.
#endif  // !PYSTL_SCAN_EMU
constexpr int kMaxAssembleCols = 16, kMaxAssembleCh = 32, kMaxAssembleD = 64;
struct AssembleArgs {
  const float* cols[kMaxAssembleCols];
  float* gcols[kMaxAssembleCols];
  int ch[kMaxAssembleCh];
  signed char inv[kMaxAssembleD];  // channel d -> its position among the read channels, -1: not read
  int n, L, D;
};
// A CTA of 256 threads owns 256 consecutive rows = 256 * W consecutive output floats: stores are fully coalesced, loads stay inside
// the rows the CTA is writing, and the (row, column) split is 32-bit arithmetic on the tile-local index (compile-time W up to 8).
template <int WT>
__global__ void __launch_bounds__(256) assemble_rows(const float* __restrict__ signal, const __grid_constant__ AssembleArgs a,
                                                     float* __restrict__ out, long long rows) {
  const int W = WT ? WT : a.n + a.L;
  for (long long r0 = (long long)blockIdx.x * 256; r0 < rows; r0 += (long long)gridDim.x * 256) {
    const int nr = rows - r0 < 256 ? (int)(rows - r0) : 256;
    const float* s = signal + r0 * a.D;
    float* o = out + r0 * W;
    for (int e = threadIdx.x; e < nr * W; e += 256) {
      const int r = e / W, w = e - r * W;
      o[e] = w < a.n ? s[r * a.D + a.ch[w]] : a.cols[w - a.n][r0 + r];
    }
  }
}
template <int DT>
__global__ void __launch_bounds__(256) scatter_rows(const float* __restrict__ grad, const __grid_constant__ AssembleArgs a,
                                                    float* __restrict__ grad_signal, long long rows) {
  const int W = a.n + a.L, D = DT ? DT : a.D;
  for (long long r0 = (long long)blockIdx.x * 256; r0 < rows; r0 += (long long)gridDim.x * 256) {
    const int nr = rows - r0 < 256 ? (int)(rows - r0) : 256;
    const float* g = grad + r0 * W;
    if (grad_signal) {
      float* s = grad_signal + r0 * D;
      for (int e = threadIdx.x; e < nr * D; e += 256) {
        const int r = e / D, i = a.inv[e - r * D];
        s[e] = i >= 0 ? g[r * W + i] : 0.f;
      }
    }
    if ((int)threadIdx.x < nr)
      for (int jc = 0; jc < a.L; ++jc)
        if (a.gcols[jc]) a.gcols[jc][r0 + threadIdx.x] = g[threadIdx.x * W + a.n + jc];
  }
}

#if !defined(PYSTL_SCAN_EMU)
static int assemble_fill(AssembleArgs& a, int32_t D, const int32_t* channels, int32_t n_channels, int32_t n_cols) {
  if (n_channels > kMaxAssembleCh || n_cols > kMaxAssembleCols || D > kMaxAssembleD)
    return scan_fail(PYSTL_E_UNSUPPORTED, "at most %d read channels of %d and %d leaf columns", kMaxAssembleCh, kMaxAssembleD, kMaxAssembleCols);
  a.n = n_channels, a.L = n_cols, a.D = D;
  for (int d = 0; d < kMaxAssembleD; ++d) a.inv[d] = -1;
  for (int i = 0; i < n_channels; ++i) {
    if (channels[i] < 0 || channels[i] >= D) return scan_fail(PYSTL_E_INVALID, "channel %d of a %d-channel signal", channels[i], D);
    if (a.inv[channels[i]] >= 0) return scan_fail(PYSTL_E_INVALID, "channel %d is listed twice", channels[i]);
    a.ch[i] = channels[i];
    a.inv[channels[i]] = (signed char)i;
  }
  return PYSTL_OK;
}
static unsigned assemble_grid(long long rows) {
  const long long blocks = (rows + 255) / 256;
  return (unsigned)(blocks < 148 * 64 ? (blocks < 1 ? 1 : blocks) : 148 * 64);  // tile-stride beyond that
}

int assemble_channels(const float* signal, int64_t rows, int32_t D, const int32_t* channels, int32_t n_channels, const float* const* cols,
                      int32_t n_cols, float* out, void* stream, char* err, size_t err_cap) {
  g_err = err;
  g_err_cap = err_cap;
  if (rows < 0 || D < 0 || n_channels < 0 || n_cols < 0 || n_channels + n_cols == 0) return scan_fail(PYSTL_E_INVALID, "bad shape");
  if ((n_channels && (!signal || !channels)) || (n_cols && !cols) || (rows && !out)) return scan_fail(PYSTL_E_INVALID, "null buffer");
  AssembleArgs a{};
  const int rc = assemble_fill(a, D, channels, n_channels, n_cols);
  if (rc) return rc;
  if (rows == 0) return PYSTL_OK;
  for (int j = 0; j < n_cols; ++j) {
    if (!cols[j]) return scan_fail(PYSTL_E_INVALID, "null column");
    a.cols[j] = cols[j];
  }
  const unsigned grid = assemble_grid(rows);
  cudaStream_t st = (cudaStream_t)stream;
  switch (n_channels + n_cols) {
#define PYSTL_ASM(WT) case WT: assemble_rows<WT><<<grid, 256, 0, st>>>(signal, a, out, (long long)rows); break;
    PYSTL_ASM(1) PYSTL_ASM(2) PYSTL_ASM(3) PYSTL_ASM(4) PYSTL_ASM(5) PYSTL_ASM(6) PYSTL_ASM(7) PYSTL_ASM(8)
#undef PYSTL_ASM
    default: assemble_rows<0><<<grid, 256, 0, st>>>(signal, a, out, (long long)rows);
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return scan_fail(PYSTL_E_CUDA, "assemble_rows: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

int scatter_channels(const float* grad, int64_t rows, int32_t D, const int32_t* channels, int32_t n_channels, float* grad_signal,
                     float* const* grad_cols, int32_t n_cols, void* stream, char* err, size_t err_cap) {
  g_err = err;
  g_err_cap = err_cap;
  if (rows < 0 || D < 0 || n_channels < 0 || n_cols < 0 || n_channels + n_cols == 0) return scan_fail(PYSTL_E_INVALID, "bad shape");
  if ((rows && !grad) || (n_channels && !channels) || (n_cols && !grad_cols)) return scan_fail(PYSTL_E_INVALID, "null buffer");
  AssembleArgs a{};
  const int rc = assemble_fill(a, D, channels, n_channels, n_cols);
  if (rc) return rc;
  if (rows == 0) return PYSTL_OK;
  for (int j = 0; j < n_cols; ++j) a.gcols[j] = grad_cols[j];
  const unsigned grid = assemble_grid(rows);
  cudaStream_t st = (cudaStream_t)stream;
  switch (D) {
#define PYSTL_SCT(DT) case DT: scatter_rows<DT><<<grid, 256, 0, st>>>(grad, a, grad_signal, (long long)rows); break;
    PYSTL_SCT(1) PYSTL_SCT(2) PYSTL_SCT(3) PYSTL_SCT(4) PYSTL_SCT(5) PYSTL_SCT(6) PYSTL_SCT(7) PYSTL_SCT(8)
#undef PYSTL_SCT
    default: scatter_rows<0><<<grid, 256, 0, st>>>(grad, a, grad_signal, (long long)rows);
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return scan_fail(PYSTL_E_CUDA, "scatter_rows: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

#endif  // !PYSTL_SCAN_EMU

}  // namespace pystl
