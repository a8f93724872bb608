// Single-source tile evaluator: the per-CTA body of the fused formula kernel.
//
// Compiled twice from this one file:
//   * by nvcc into the sm_100a kernels of libpystl_b200.so (DevExec: one CUDA
//     thread per lane, __syncthreads, warp shuffles), and
//   * by g++ into tests/_emu/libpystl_emu.so (HostExec: the CTA's threads are
//     run one after another between barriers).  That library is TEST
//     INFRASTRUCTURE: it lets the CPU-only test tier check the kernel's index
//     arithmetic against the oracle.  Nothing in pystl_b200/ loads it.
//
// Data layout inside a tile of `lext = 8 * nt` samples:
//   plane   one float per sample; thread `tid` owns samples [8 tid, 8 tid + 8)
//           as two float4.  The two float4 of a thread are stored swapped when
//           bit 2 of tid is set (pf4), which makes every LDS.128 / STS.128 of a
//           quarter-warp hit 8 distinct 16-byte bank groups.
//   value   1 plane (2 for cumulative: pos, neg) per live formula node
//   scratch planes shared by the instruction that is executing
//
// Reference semantics restated here (file:line under /root/reference/pystl):
//   classic  min / max                      semantics/classic.py:25-37
//   smooth   -/+ tau*logsumexp(-/+ v/tau)   semantics/smooth.py:18-43
//   cumul.   (pos,neg): min / max / SUM     semantics/cumulative.py:35-72
//   agm      arithmetic / geometric mean    semantics/agm.py:60-88
//   dgmsr    gmsr_and / gmsr_or             semantics/_dgmsr.py:66-119
//   windows  [t+a, min(t+b, T-1)]           api.py:49-54
//   until    OR_i AND(AND_{k<=i} l, r_i)    api.py:284-302 + each semantics' temporal_until
#pragma once
#include <math.h>
#include <stdint.h>

#include "program.h"

#if defined(__CUDACC__)
#define PYSTL_HD __host__ __device__ __forceinline__
#else
#define PYSTL_HD inline
#endif

#define PYSTL_FOR_THREADS(ex, tid) for (int tid = (ex).t0(), tid##_end_ = (ex).t1(); tid < tid##_end_; ++tid)

namespace pystl {

struct alignas(16) F4 {
  float v[4];
};

PYSTL_HD int pf4(int f) { return f ^ ((f >> 3) & 1); }          // physical float4 index
PYSTL_HD int pidx(int p) { return p ^ (((p >> 5) & 1) << 2); }  // physical scalar index

PYSTL_HD float f_inf() { return INFINITY; }
PYSTL_HD float f_nan() { return NAN; }

PYSTL_HD float fast_ex2(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return exp2f(x);
#endif
}
PYSTL_HD float fast_lg2(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return log2f(x);
#endif
}

// ------------------------------------------------------------------ run-time parameters
struct RP {
  float k;         // smooth: log2(e) / tau
  float invk;      // smooth: tau * ln 2
  float shift;     // smooth: block-wide max of the exponent, set per instruction
  float eps;       // dgmsr
  float sqrt_eps;  // dgmsr
  float ln_eps;    // dgmsr
  float eps_p;     // dgmsr eps^p
  float inv2p;     // dgmsr 1 / (2p)
  int p;
};

PYSTL_HD RP make_rp(const TileProgram& P) {
  RP r;
  const float tau = P.tau > 0.f ? P.tau : 1.f;
  r.k = 1.4426950408889634f / tau;
  r.invk = tau * 0.6931471805599453f;
  r.shift = 0.f;
  r.eps = P.eps;
  r.sqrt_eps = sqrtf(P.eps);
  r.ln_eps = logf(P.eps);
  r.p = P.p < 1 ? 1 : P.p;
  r.eps_p = powf(P.eps, (float)r.p);
  r.inv2p = 0.5f / (float)r.p;
  return r;
}

// ------------------------------------------------------------------ values and accumulators
template <int SEM>
struct Tr {
  static constexpr int VW = (SEM == SEM_CUMULATIVE) ? 2 : 1;                       // words per node value
  static constexpr int AW = (SEM == SEM_CLASSIC || SEM == SEM_SMOOTH) ? 1 : 2;     // words per accumulator
};

template <int N>
struct Vec {
  float x[N];
};

template <int SEM>
PYSTL_HD Vec<Tr<SEM>::VW> negate(Vec<Tr<SEM>::VW> v) {
  Vec<Tr<SEM>::VW> r;
  if constexpr (SEM == SEM_CUMULATIVE) {  // cumulative.py:32-33
    r.x[0] = -v.x[1];
    r.x[1] = -v.x[0];
  } else {
    r.x[0] = -v.x[0];
  }
  return r;
}

// Reducer: ident / lift (value -> accumulator) / comb (earlier, later) / fin (accumulator -> value).
template <int SEM, int MODE>
struct Red;

template <int MODE>
struct Red<SEM_CLASSIC, MODE> {
  using V = Vec<1>;
  using A = Vec<1>;
  static PYSTL_HD A ident() { return A{{MODE == MODE_AND ? f_inf() : -f_inf()}}; }
  static PYSTL_HD A lift(V v, const RP&, float) { return A{{v.x[0]}}; }
  static PYSTL_HD A comb(A a, A b) { return A{{MODE == MODE_AND ? fminf(a.x[0], b.x[0]) : fmaxf(a.x[0], b.x[0])}}; }
  static PYSTL_HD V fin(A a, float, const RP&) { return V{{a.x[0]}}; }
};

template <int MODE>
struct Red<SEM_SMOOTH, MODE> {
  using V = Vec<1>;
  using A = Vec<1>;
  static PYSTL_HD float expo(V v, const RP& rp) { return (MODE == MODE_AND ? -v.x[0] : v.x[0]) * rp.k; }
  static PYSTL_HD A ident() { return A{{0.f}}; }
  static PYSTL_HD A lift(V v, const RP& rp, float) { return A{{fast_ex2(expo(v, rp) - rp.shift)}}; }
  static PYSTL_HD A comb(A a, A b) { return A{{a.x[0] + b.x[0]}}; }
  static PYSTL_HD V fin(A a, float, const RP& rp) {
    const float l = (fast_lg2(a.x[0]) + rp.shift) * rp.invk;
    return V{{MODE == MODE_AND ? -l : l}};
  }
};

template <int MODE>
struct Red<SEM_CUMULATIVE, MODE> {
  using V = Vec<2>;
  using A = Vec<2>;
  static PYSTL_HD A ident() {
    const float i = MODE == MODE_AND ? f_inf() : (MODE == MODE_OR ? -f_inf() : 0.f);
    return A{{i, i}};
  }
  static PYSTL_HD A lift(V v, const RP&, float) { return A{{v.x[0], v.x[1]}}; }
  static PYSTL_HD A comb(A a, A b) {
    if constexpr (MODE == MODE_AND) return A{{fminf(a.x[0], b.x[0]), fminf(a.x[1], b.x[1])}};
    if constexpr (MODE == MODE_OR) return A{{fmaxf(a.x[0], b.x[0]), fmaxf(a.x[1], b.x[1])}};
    return A{{a.x[0] + b.x[0], a.x[1] + b.x[1]}};
  }
  static PYSTL_HD V fin(A a, float, const RP&) { return V{{a.x[0], a.x[1]}}; }
};

// AGM accumulator: x[0] = sum w*clip(v), x[1] = sum w*log1p(+-v), or +inf once an
// entry takes the arithmetic branch (log1p of a finite value is finite, so the
// flag cannot be produced by accident).
template <int MODE>
struct Red<SEM_AGM, MODE> {
  using V = Vec<1>;
  using A = Vec<2>;
  static PYSTL_HD A ident() { return A{{0.f, 0.f}}; }
  static PYSTL_HD A lift(V v, const RP&, float w) {
    const float x = v.x[0];
    if constexpr (MODE == MODE_AND) {  // agm.py:71-73
      return (x <= 0.f) ? A{{w * x, f_inf()}} : A{{0.f, w * log1pf(x)}};
    } else {  // agm.py:86-88
      return (x > 0.f) ? A{{w * x, f_inf()}} : A{{0.f, w * log1pf(-x)}};
    }
  }
  static PYSTL_HD A comb(A a, A b) { return A{{a.x[0] + b.x[0], a.x[1] + b.x[1]}}; }
  static PYSTL_HD V fin(A a, float sw, const RP&) {
    const float inv = 1.f / sw;
    if (a.x[1] == f_inf()) return V{{a.x[0] * inv}};
    const float g = expm1f(a.x[1] * inv);
    return V{{MODE == MODE_AND ? g : -g}};
  }
};

// D-GMSR accumulator: x[0] = sum_{v<=0} w v^(2p), x[1] = sum_{v>0} 2 w ln v or +inf (flag).
// OR is -AND(-v) (_dgmsr.py:110-119).
template <int MODE>
struct Red<SEM_DGMSR, MODE> {
  using V = Vec<1>;
  using A = Vec<2>;
  static PYSTL_HD A ident() { return A{{0.f, 0.f}}; }
  static PYSTL_HD A lift(V v, const RP& rp, float w) {
    const float x = MODE == MODE_AND ? v.x[0] : -v.x[0];
    if (x <= 0.f) {
      const float x2 = x * x;
      return A{{w * (rp.p == 1 ? x2 : powf(x2, (float)rp.p)), f_inf()}};
    }
    return A{{0.f, 2.f * w * logf(x)}};
  }
  static PYSTL_HD A comb(A a, A b) { return A{{a.x[0] + b.x[0], a.x[1] + b.x[1]}}; }
  static PYSTL_HD V fin(A a, float sw, const RP& rp) {
    float h;
    if (a.x[1] == f_inf()) {  // _dgmsr.py:83-87
      const float m = rp.eps_p + a.x[0] / sw;
      h = rp.sqrt_eps - (rp.p == 1 ? sqrtf(m) : powf(m, rp.inv2p));
    } else {  // _dgmsr.py:96-102, in the log domain: ln m0 = logaddexp(sw ln eps, sum 2w ln v) / sw
      const float u = sw * rp.ln_eps, s = a.x[1];
      const float hi = fmaxf(u, s), lo = fminf(u, s);
      const float lm0 = (hi + log1pf(expf(lo - hi))) / sw;
      h = expf(0.5f * lm0) - rp.sqrt_eps;
    }
    return V{{MODE == MODE_AND ? h : -h}};
  }
};

// ------------------------------------------------------------------ tile context
struct Ctx {
  float* smem;
  TileGeom g;
  const TileProgram* P;
  int64_t e0;  // trace time of tile sample 0
  int64_t T;
  RP rp;
  PYSTL_HD float* plane(int slot, int w, int vw) const { return smem + g.off_planes + (slot * vw + w) * g.lext; }
  PYSTL_HD float* scr(int i) const { return smem + g.off_scratch + i * g.lext; }
  PYSTL_HD float* chunk(int i) const { return smem + g.off_chunk + i * g.nt; }
  PYSTL_HD float* red() const { return smem + g.off_red; }
  // number of tile samples that exist in the trace
  PYSTL_HD int n_live() const {
    const int64_t r = T - e0;
    return r < (int64_t)g.lext ? (int)r : g.lext;
  }
};

// own chunk of a plane <-> registers
PYSTL_HD void load8(const float* plane, int tid, float (&v)[8]) {
  const F4* P4 = reinterpret_cast<const F4*>(plane);
  const int s = (tid >> 2) & 1;
  const F4 lo = P4[2 * tid + s], hi = P4[2 * tid + (s ^ 1)];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = lo.v[k];
    v[4 + k] = hi.v[k];
  }
}
PYSTL_HD void store8(float* plane, int tid, const float (&v)[8]) {
  F4* P4 = reinterpret_cast<F4*>(plane);
  const int s = (tid >> 2) & 1;
  F4 lo, hi;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    lo.v[k] = v[k];
    hi.v[k] = v[4 + k];
  }
  P4[2 * tid + s] = lo;
  P4[2 * tid + (s ^ 1)] = hi;
}

template <int SEM>
PYSTL_HD void load_val8(const Ctx& c, int slot, bool neg, int tid, Vec<Tr<SEM>::VW> (&v)[8]) {
  constexpr int VW = Tr<SEM>::VW;
  float t[VW][8];
#pragma unroll
  for (int w = 0; w < VW; ++w) load8(c.plane(slot, w, VW), tid, t[w]);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int w = 0; w < VW; ++w) v[i].x[w] = t[w][i];
    if (neg) v[i] = negate<SEM>(v[i]);
  }
}
template <int SEM>
PYSTL_HD void store_val8(const Ctx& c, int slot, int tid, const Vec<Tr<SEM>::VW> (&v)[8]) {
  constexpr int VW = Tr<SEM>::VW;
#pragma unroll
  for (int w = 0; w < VW; ++w) {
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = v[i].x[w];
    store8(c.plane(slot, w, VW), tid, t);
  }
}

// run of 8 consecutive logical samples starting at `start` (any alignment, any sign)
// from a scratch plane; positions outside [0, lext) read as `fill`.
PYSTL_HD void load_run8(const float* plane, int lext, int start, float fill, float (&v)[8]) {
  const F4* P4 = reinterpret_cast<const F4*>(plane);
  const int f0 = start >> 2;  // arithmetic shift: floor
  const int r = start & 3;
  const int nf = lext >> 2;
  float t[12];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int f = f0 + j;
    if (f >= 0 && f < nf) {
      const F4 q = P4[pf4(f)];
#pragma unroll
      for (int k = 0; k < 4; ++k) t[4 * j + k] = q.v[k];
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) t[4 * j + k] = fill;
    }
  }
  // r is uniform across the CTA: a 4-way switch keeps every register index static
  switch (r) {
    case 0:
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = t[i];
      break;
    case 1:
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = t[i + 1];
      break;
    case 2:
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = t[i + 2];
      break;
    default:
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = t[i + 3];
      break;
  }
}

// ------------------------------------------------------------------ predicates
// Strided ownership (sample p = tid + nt*i) so that the raw [sample][channel]
// rows are read without bank conflicts; results go to the predicate planes.
template <int SEM, class EX>
PYSTL_HD void eval_predicates(EX& ex, const Ctx& c, const float* raw) {
  constexpr int VW = Tr<SEM>::VW;
  const TileProgram& P = *c.P;
  const int D = c.g.D, nlive = c.n_live();
  PYSTL_FOR_THREADS(ex, tid) {
    for (int i = 0; i < kChunk; ++i) {
      const int p = tid + c.g.nt * i;
      if (p >= nlive) continue;
      const float* row = raw + (int64_t)p * D;
      for (int q = 0; q < P.n_preds; ++q) {
        const TPred& pr = P.preds[q];
        float acc;
        if (SEM == SEM_CLASSIC && pr.f64) {
          // the reference evaluates predicates in float64 (classic.py:20); same operations, same
          // order as formula.affine_predicate, one final rounding: bit-exact after the cast
          double a = 0.0;
          for (int k = 0; k < pr.n_terms; ++k) a = fma((double)pr.coeff[k], (double)row[pr.ch[k]], a);
          acc = (float)(a + (double)pr.bias);
        } else {
          acc = 0.f;  // 0 + (+-x) is exact, so single-term predicates round once
          for (int k = 0; k < pr.n_terms; ++k) acc = fmaf(pr.coeff[k], row[pr.ch[k]], acc);
          acc += pr.bias;
        }
        const int pp = pidx(p);
        if constexpr (SEM == SEM_CUMULATIVE) {  // cumulative.py:29-30
          c.plane(pr.dst, 0, VW)[pp] = fmaxf(0.f, acc);
          c.plane(pr.dst, 1, VW)[pp] = fminf(0.f, acc);
        } else {
          c.plane(pr.dst, 0, VW)[pp] = acc;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ n-ary And / Or
template <int SEM, int MODE, class EX>
PYSTL_HD void op_nary(EX& ex, const Ctx& c, const TInstr& in) {
  using R = Red<SEM, MODE>;
  using V = typename R::V;
  using A = typename R::A;
  PYSTL_FOR_THREADS(ex, tid) {
    RP rp = c.rp;
    A acc[8];
    if constexpr (SEM == SEM_SMOOTH) {
      // exact max-shifted logsumexp per sample (smooth.py:18-22): pass 1 finds the max exponent
      float m[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = -f_inf();
      for (int j = 0; j < in.nsrc; ++j) {
        V v[8];
        load_val8<SEM>(c, in.src[j], (in.negmask >> j) & 1, tid, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], R::expo(v[i], rp));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = R::ident();
      for (int j = 0; j < in.nsrc; ++j) {
        V v[8];
        load_val8<SEM>(c, in.src[j], (in.negmask >> j) & 1, tid, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i].x[0] += fast_ex2(R::expo(v[i], rp) - m[i]);
      }
      V o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        rp.shift = m[i];
        o[i] = R::fin(acc[i], in.sw, rp);
      }
      store_val8<SEM>(c, in.dst, tid, o);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = R::ident();
      for (int j = 0; j < in.nsrc; ++j) {
        V v[8];
        load_val8<SEM>(c, in.src[j], (in.negmask >> j) & 1, tid, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = R::comb(acc[i], R::lift(v[i], rp, in.w[j]));
      }
      V o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = R::fin(acc[i], in.sw, rp);
      store_val8<SEM>(c, in.dst, tid, o);
    }
  }
}

// number of samples of the clipped window [t+a, min(t+b, T-1)] (api.py:49-54)
PYSTL_HD int window_count(int64_t t, int a, int b, int64_t T) {
  const int64_t lo = t + a, hi = (t + b < T - 1) ? t + b : T - 1;
  const int64_t lo2 = lo < 0 ? 0 : lo;
  return hi >= lo2 ? (int)(hi - lo2 + 1) : 0;
}

// smooth only: block-wide max exponent of the operand over the live samples
template <int SEM, int MODE, class EX>
PYSTL_HD float smooth_block_shift(EX& ex, const Ctx& c, int slot, bool neg) {
  using R = Red<SEM, MODE>;
  using V = typename R::V;
  const int nlive = c.n_live();
  PYSTL_FOR_THREADS(ex, tid) {
    V v[8];
    load_val8<SEM>(c, slot, neg, tid, v);
    float m = -f_inf();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (8 * tid + i < nlive) m = fmaxf(m, R::expo(v[i], c.rp));
    ex.put(tid, m);
  }
  return ex.all_max(c.red());
}

// smooth only: exact logsumexp of one window straight from the operand plane (rare path,
// taken when the block-shifted sum lost its precision to underflow)
template <int SEM, int MODE>
PYSTL_HD Vec<1> smooth_exact_window(const Ctx& c, int slot, bool neg, int p, int a, int b) {
  using R = Red<SEM, MODE>;
  const float* pl = c.plane(slot, 0, 1);
  const int nlive = c.n_live();
  int lo = p + a, hi = p + b;
  if (lo < 0) lo = 0;
  if (hi > nlive - 1) hi = nlive - 1;
  float m = -f_inf();
  for (int q = lo; q <= hi; ++q) {
    Vec<1> v{{neg ? -pl[pidx(q)] : pl[pidx(q)]}};
    m = fmaxf(m, R::expo(v, c.rp));
  }
  float s = 0.f;
  for (int q = lo; q <= hi; ++q) {
    Vec<1> v{{neg ? -pl[pidx(q)] : pl[pidx(q)]}};
    s += fast_ex2(R::expo(v, c.rp) - m);
  }
  RP rp = c.rp;
  rp.shift = m;
  return R::fin(Vec<1>{{s}}, 1.f, rp);
}

constexpr float kSmoothTiny = 7.8886090522101181e-31f;  // 2^-100

template <int SEM, int MODE>
PYSTL_HD typename Red<SEM, MODE>::V finish_window(const Ctx& c, const TInstr& in, const RP& rp,
                                                  typename Red<SEM, MODE>::A acc, int p) {
  using R = Red<SEM, MODE>;
  const int n = window_count(c.e0 + p, in.a, in.b, c.T);
  typename R::V o;
  if (n <= 0) {
#pragma unroll
    for (int w = 0; w < Tr<SEM>::VW; ++w) o.x[w] = f_nan();
    return o;
  }
  if constexpr (SEM == SEM_SMOOTH) {
    if (!(acc.x[0] >= kSmoothTiny)) return smooth_exact_window<SEM, MODE>(c, in.src[0], in.negmask & 1, p, in.a, in.b);
  }
  return R::fin(acc, (float)n, rp);
}

// lift the own chunk of the operand, treating samples past the horizon as the identity
template <int SEM, int MODE>
PYSTL_HD void lift_own8(const Ctx& c, const TInstr& in, const RP& rp, int tid, typename Red<SEM, MODE>::A (&x)[8]) {
  using R = Red<SEM, MODE>;
  typename R::V v[8];
  load_val8<SEM>(c, in.src[0], in.negmask & 1, tid, v);
  const int nlive = c.n_live();
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (8 * tid + i < nlive) ? R::lift(v[i], rp, 1.f) : R::ident();
}

// ------------------------------------------------------------------ window, direct (W <= 8)
template <int SEM, int MODE, class EX>
PYSTL_HD void op_win_direct(EX& ex, const Ctx& c, const TInstr& in) {
  using R = Red<SEM, MODE>;
  using A = typename R::A;
  using V = typename R::V;
  constexpr int AW = Tr<SEM>::AW;
  RP rp = c.rp;
  if constexpr (SEM == SEM_SMOOTH) rp.shift = smooth_block_shift<SEM, MODE>(ex, c, in.src[0], in.negmask & 1);
  PYSTL_FOR_THREADS(ex, tid) {
    A x[8];
    lift_own8<SEM, MODE>(c, in, rp, tid, x);
#pragma unroll
    for (int w = 0; w < AW; ++w) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = x[i].x[w];
      store8(c.scr(w), tid, t);
    }
  }
  ex.sync();
  PYSTL_FOR_THREADS(ex, tid) {
    const int p0 = 8 * tid, W = in.b - in.a + 1;
    const A id = R::ident();
    A acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = id;
    const int nf = c.g.lext >> 2;
    const int f_lo = (p0 + in.a) >> 2, f_hi = (p0 + 7 + in.b) >> 2;
    for (int f = f_lo; f <= f_hi; ++f) {
      if (f < 0 || f >= nf) continue;  // outside the tile: identity
      F4 q[AW];
#pragma unroll
      for (int w = 0; w < AW; ++w) q[w] = reinterpret_cast<const F4*>(c.scr(w))[pf4(f)];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        A e;
#pragma unroll
        for (int w = 0; w < AW; ++w) e.x[w] = q[w].v[k];
        const int d = 4 * f + k - p0 - in.a;  // window slot of output 0 this element fills
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (d - i >= 0 && d - i < W) acc[i] = R::comb(acc[i], e);
      }
    }
    V o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = finish_window<SEM, MODE>(c, in, rp, acc[i], p0 + i);
    store_val8<SEM>(c, in.dst, tid, o);
  }
  ex.sync();
}

// aggregate of chunk `ch`, identity outside the tile
template <int AW>
PYSTL_HD Vec<AW> load_agg(const Ctx& c, int ch, int nt, Vec<AW> id) {
  Vec<AW> r = id;
  if (ch >= 0 && ch < nt) {
#pragma unroll
    for (int w = 0; w < AW; ++w) r.x[w] = c.chunk(w)[ch];
  }
  return r;
}

// ------------------------------------------------------------------ window, van Herk / Gil-Werman (W >= 9)
// Chunk = the 8 samples of one thread.  Phase 1: in-register suffix and prefix
// scans of the own chunk -> scratch planes SUF, PRE and the chunk aggregate AGG.
// Phase 2: out[t] = SUF[t+a] (+) AGG[whole chunks strictly inside] (+) PRE[t+b].
// The pieces are disjoint, so (+) only has to be associative (min, max, +).
template <int SEM, int MODE, class EX>
PYSTL_HD void op_win_vh(EX& ex, const Ctx& c, const TInstr& in) {
  using R = Red<SEM, MODE>;
  using A = typename R::A;
  using V = typename R::V;
  constexpr int AW = Tr<SEM>::AW;
  RP rp = c.rp;
  if constexpr (SEM == SEM_SMOOTH) rp.shift = smooth_block_shift<SEM, MODE>(ex, c, in.src[0], in.negmask & 1);
  PYSTL_FOR_THREADS(ex, tid) {
    A x[8], suf[8], pre[8];
    lift_own8<SEM, MODE>(c, in, rp, tid, x);
    pre[0] = x[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) pre[i] = R::comb(pre[i - 1], x[i]);
    suf[7] = x[7];
#pragma unroll
    for (int i = 6; i >= 0; --i) suf[i] = R::comb(x[i], suf[i + 1]);
#pragma unroll
    for (int w = 0; w < AW; ++w) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = suf[i].x[w];
      store8(c.scr(w), tid, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = pre[i].x[w];
      store8(c.scr(AW + w), tid, t);
      c.chunk(w)[tid] = pre[7].x[w];
    }
  }
  ex.sync();
  PYSTL_FOR_THREADS(ex, tid) {
    const int nt = c.g.nt, lext = c.g.lext;
    const A id = R::ident();
    const int qa = in.a >> 3, ra = in.a & 7, qb = in.b >> 3, rb = in.b & 7;
    float sufv[AW][8], prev[AW][8];
#pragma unroll
    for (int w = 0; w < AW; ++w) {
      load_run8(c.scr(w), lext, 8 * tid + in.a, id.x[w], sufv[w]);
      load_run8(c.scr(AW + w), lext, 8 * tid + in.b, id.x[w], prev[w]);
    }
    const A al = load_agg<AW>(c, tid + qa + 1, nt, id), ar = load_agg<AW>(c, tid + qb, nt, id);
    A core = id;
    for (int ch = tid + qa + 2; ch <= tid + qb - 1; ++ch) core = R::comb(core, load_agg<AW>(c, ch, nt, id));
    const bool two = qb >= qa + 2;
    V o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      A s, e;
#pragma unroll
      for (int w = 0; w < AW; ++w) {
        s.x[w] = sufv[w][i];
        e.x[w] = prev[w][i];
      }
      const int cs = (ra + i) >> 3, ce = (rb + i) >> 3;  // 0 or 1: did start / end spill into the next chunk
      const bool inc_l = (cs == 0) && (qb + ce - 1 >= qa + 1);
      const bool inc_r = (ce == 1) && two;
      A acc = s;
      if (inc_l) acc = R::comb(acc, al);
      acc = R::comb(acc, core);
      if (inc_r) acc = R::comb(acc, ar);
      acc = R::comb(acc, e);
      o[i] = finish_window<SEM, MODE>(c, in, rp, acc, 8 * tid + i);
    }
    store_val8<SEM>(c, in.dst, tid, o);
  }
  ex.sync();
}

// ------------------------------------------------------------------ open-ended window [a, horizon]
// Only lowered when the tile reaches the end of the trace.  Suffix scan of the
// whole tile (chunk-local in registers + a block scan of the chunk aggregates),
// then a shift by `a`.
template <int SEM, int MODE>
struct RedOp {
  using R = Red<SEM, MODE>;
  using A = typename R::A;
  static PYSTL_HD A ident() { return R::ident(); }
  static PYSTL_HD A comb(A a, A b) { return R::comb(a, b); }
};

template <int SEM, int MODE, class EX>
PYSTL_HD void op_win_suffix(EX& ex, const Ctx& c, const TInstr& in) {
  using R = Red<SEM, MODE>;
  using A = typename R::A;
  using V = typename R::V;
  constexpr int AW = Tr<SEM>::AW;
  RP rp = c.rp;
  if constexpr (SEM == SEM_SMOOTH) rp.shift = smooth_block_shift<SEM, MODE>(ex, c, in.src[0], in.negmask & 1);
  PYSTL_FOR_THREADS(ex, tid) {
    A x[8];
    lift_own8<SEM, MODE>(c, in, rp, tid, x);
    A agg = x[7];
#pragma unroll
    for (int i = 6; i >= 0; --i) agg = R::comb(x[i], agg);
#pragma unroll
    for (int w = 0; w < AW; ++w) c.chunk(w)[tid] = agg.x[w];
  }
  ex.sync();
  ex.template excl_suffix_scan<AW, RedOp<SEM, MODE>>(c.chunk(0), c.g.nt, c.red());
  PYSTL_FOR_THREADS(ex, tid) {
    A x[8], suf[8];
    lift_own8<SEM, MODE>(c, in, rp, tid, x);
    A tail;
#pragma unroll
    for (int w = 0; w < AW; ++w) tail.x[w] = c.chunk(w)[tid];
    suf[7] = R::comb(x[7], tail);
#pragma unroll
    for (int i = 6; i >= 0; --i) suf[i] = R::comb(x[i], suf[i + 1]);
#pragma unroll
    for (int w = 0; w < AW; ++w) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = suf[i].x[w];
      store8(c.scr(w), tid, t);
    }
  }
  ex.sync();
  PYSTL_FOR_THREADS(ex, tid) {
    const A id = R::ident();
    float sv[AW][8];
#pragma unroll
    for (int w = 0; w < AW; ++w) load_run8(c.scr(w), c.g.lext, 8 * tid + in.a, id.x[w], sv[w]);
    V o[8];
    TInstr open = in;
    open.b = 0x3fffffff;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      A acc;
#pragma unroll
      for (int w = 0; w < AW; ++w) acc.x[w] = sv[w][i];
      o[i] = finish_window<SEM, MODE>(c, open, rp, acc, 8 * tid + i);
    }
    store_val8<SEM>(c, in.dst, tid, o);
  }
  ex.sync();
}

// ------------------------------------------------------------------ classic unbounded until (clamp scan)
// W(s) = min(l_s, max(r_s, W(s+1))) is the composition of f_s(x) = min(A, max(B, x)),
// which is closed under composition, so the whole tile is one
// suffix scan.  Starting the recurrence at the tile's last sample is exact for every
// output whose bounded window lies inside the tile ("Classic until").
struct ClampOp {
  using A = Vec<2>;
  static PYSTL_HD A ident() { return A{{f_inf(), -f_inf()}}; }
  static PYSTL_HD A comb(A f, A g) {  // f after g:  x -> f(g(x)), f is the EARLIER sample
    return A{{fminf(f.x[0], fmaxf(f.x[1], g.x[0])), fmaxf(f.x[1], g.x[1])}};
  }
};

PYSTL_HD void clamp_load_lr(const Ctx& c, const TInstr& in, int tid, int nlive, ClampOp::A (&x)[8]) {
  Vec<1> l[8], r[8];
  load_val8<SEM_CLASSIC>(c, in.src[0], in.negmask & 1, tid, l);
  load_val8<SEM_CLASSIC>(c, in.src[1], (in.negmask >> 1) & 1, tid, r);
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (8 * tid + i < nlive) ? ClampOp::A{{l[i].x[0], r[i].x[0]}} : ClampOp::ident();
}

template <class EX>
PYSTL_HD void op_clamp_scan(EX& ex, const Ctx& c, const TInstr& in) {
  using A = ClampOp::A;
  const int nlive = c.n_live();
  PYSTL_FOR_THREADS(ex, tid) {
    A x[8];
    clamp_load_lr(c, in, tid, nlive, x);
    A agg = x[7];
#pragma unroll
    for (int i = 6; i >= 0; --i) agg = ClampOp::comb(x[i], agg);
    c.chunk(0)[tid] = agg.x[0];
    c.chunk(1)[tid] = agg.x[1];
  }
  ex.sync();
  ex.template excl_suffix_scan<2, ClampOp>(c.chunk(0), c.g.nt, c.red());
  PYSTL_FOR_THREADS(ex, tid) {
    A x[8];
    clamp_load_lr(c, in, tid, nlive, x);
    A tail{{c.chunk(0)[tid], c.chunk(1)[tid]}};
    Vec<1> o[8];
    A run = tail;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      run = ClampOp::comb(x[i], run);
      o[i].x[0] = fminf(run.x[0], run.x[1]);  // f(-inf) = min(A, B)
    }
    store_val8<SEM_CLASSIC>(c, in.dst, tid, o);
  }
  ex.sync();
}

// ------------------------------------------------------------------ until, O(b) per sample
// Per-semantics running state of one output while its window slides to the right.
template <int SEM>
struct UntilState;

template <>
struct UntilState<SEM_CLASSIC> {  // classic.py:63-67
  float pm, best;
  PYSTL_HD void init() { pm = f_inf(); best = -f_inf(); }
  PYSTL_HD void step(Vec<1> l, Vec<1> r, bool cand, const RP&, const TInstr&) {
    pm = fminf(pm, l.x[0]);
    if (cand) best = fmaxf(best, fminf(pm, r.x[0]));
  }
  PYSTL_HD Vec<1> done(int, const RP&) { return Vec<1>{{best}}; }
};

template <>
struct UntilState<SEM_CUMULATIVE> {  // cumulative.py:96-102
  float pm0, pm1, s0, s1;
  PYSTL_HD void init() { pm0 = pm1 = f_inf(); s0 = s1 = 0.f; }
  PYSTL_HD void step(Vec<2> l, Vec<2> r, bool cand, const RP&, const TInstr&) {
    pm0 = fminf(pm0, l.x[0]);
    pm1 = fminf(pm1, l.x[1]);
    if (cand) {
      s0 += fminf(r.x[0], pm0);
      s1 += fminf(r.x[1], pm1);
    }
  }
  PYSTL_HD Vec<2> done(int, const RP&) { return Vec<2>{{s0, s1}}; }
};

// smooth.py:95-102 through the closed form of SURVEY.md appendix A4:
//   softmin(softmin(l_0..l_i), r_i) = softmin(l_0..l_i, r_i), hence
//   U = tau * log sum_i 1 / (S_i + e^{-r_i/tau}),  S_i = sum_{k<=i} e^{-l_k/tau};
// exponents are shifted by the block-wide maximum rp.shift.
template <>
struct UntilState<SEM_SMOOTH> {
  float S, acc;
  PYSTL_HD void init() { S = 0.f; acc = 0.f; }
  PYSTL_HD void step(Vec<1> l, Vec<1> r, bool cand, const RP& rp, const TInstr&) {
    S += fast_ex2(-l.x[0] * rp.k - rp.shift);
    if (cand) acc += 1.f / (S + fast_ex2(-r.x[0] * rp.k - rp.shift));
  }
  PYSTL_HD Vec<1> done(int, const RP& rp) { return Vec<1>{{(fast_lg2(acc) - rp.shift) * rp.invk}}; }
};

template <>
struct UntilState<SEM_AGM> {  // agm.py:112-130 (uniform weights_left / weights_right)
  Vec<2> pre, orr;
  int k;
  PYSTL_HD void init() { pre = Red<SEM_AGM, MODE_AND>::ident(); orr = Red<SEM_AGM, MODE_OR>::ident(); k = 0; }
  PYSTL_HD void step(Vec<1> l, Vec<1> r, bool cand, const RP& rp, const TInstr& in) {
    using RA = Red<SEM_AGM, MODE_AND>;
    using RO = Red<SEM_AGM, MODE_OR>;
    pre = RA::comb(pre, RA::lift(l, rp, 1.f));
    ++k;
    if (cand) {
      const Vec<1> y = RA::fin(pre, (float)k, rp);
      const Vec<2> pr = RA::comb(RA::lift(y, rp, in.w[0]), RA::lift(r, rp, in.w[1]));
      orr = RO::comb(orr, RO::lift(RA::fin(pr, in.w[0] + in.w[1], rp), rp, 1.f));
    }
  }
  PYSTL_HD Vec<1> done(int ncand, const RP& rp) { return Red<SEM_AGM, MODE_OR>::fin(orr, (float)ncand, rp); }
};

template <>
struct UntilState<SEM_DGMSR> {  // dgmsr.py:88-118 (uniform weights_left / weights_right)
  Vec<2> pre, orr;
  int k;
  PYSTL_HD void init() { pre = Red<SEM_DGMSR, MODE_AND>::ident(); orr = Red<SEM_DGMSR, MODE_OR>::ident(); k = 0; }
  PYSTL_HD void step(Vec<1> l, Vec<1> r, bool cand, const RP& rp, const TInstr& in) {
    using RA = Red<SEM_DGMSR, MODE_AND>;
    using RO = Red<SEM_DGMSR, MODE_OR>;
    pre = RA::comb(pre, RA::lift(l, rp, 1.f));
    ++k;
    if (cand) {
      const Vec<1> y = RA::fin(pre, (float)k, rp);
      const Vec<2> pr = RA::comb(RA::lift(y, rp, in.w[0]), RA::lift(r, rp, in.w[1]));
      orr = RO::comb(orr, RO::lift(RA::fin(pr, in.w[0] + in.w[1], rp), rp, 1.f));
    }
  }
  PYSTL_HD Vec<1> done(int ncand, const RP& rp) { return Red<SEM_DGMSR, MODE_OR>::fin(orr, (float)ncand, rp); }
};

template <int SEM, class EX>
PYSTL_HD void op_until(EX& ex, const Ctx& c, const TInstr& in) {
  constexpr int VW = Tr<SEM>::VW;
  using V = Vec<VW>;
  const int nlive = c.n_live();
  RP rp = c.rp;
  if constexpr (SEM == SEM_SMOOTH) {
    // shift = max over both operands of the exponent -v/tau
    const float ml = smooth_block_shift<SEM, MODE_AND>(ex, c, in.src[0], in.negmask & 1);
    const float mr = smooth_block_shift<SEM, MODE_AND>(ex, c, in.src[1], (in.negmask >> 1) & 1);
    rp.shift = fmaxf(ml, mr);
  }
  // stage l and r (negations applied) in scratch planes so that phase 2 reads plain runs
  PYSTL_FOR_THREADS(ex, tid) {
    V l[8], r[8];
    load_val8<SEM>(c, in.src[0], in.negmask & 1, tid, l);
    load_val8<SEM>(c, in.src[1], (in.negmask >> 1) & 1, tid, r);
#pragma unroll
    for (int w = 0; w < VW; ++w) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = l[i].x[w];
      store8(c.scr(w), tid, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = r[i].x[w];
      store8(c.scr(VW + w), tid, t);
    }
  }
  ex.sync();
  PYSTL_FOR_THREADS(ex, tid) {
    const int p0 = 8 * tid;
    const int b = in.unbounded ? c.g.lext : in.b;
    UntilState<SEM> st[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) st[i].init();
    int q_hi = p0 + 7 + b;
    if (q_hi > nlive - 1) q_hi = nlive - 1;
    for (int f = p0 >> 2; f <= (q_hi >> 2); ++f) {
      F4 ql[VW], qr[VW];
#pragma unroll
      for (int w = 0; w < VW; ++w) {
        ql[w] = reinterpret_cast<const F4*>(c.scr(w))[pf4(f)];
        qr[w] = reinterpret_cast<const F4*>(c.scr(VW + w))[pf4(f)];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int q = 4 * f + k;
        if (q > q_hi) break;
        V l, r;
#pragma unroll
        for (int w = 0; w < VW; ++w) {
          l.x[w] = ql[w].v[k];
          r.x[w] = qr[w].v[k];
        }
        const int d = q - p0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int kk = d - i;  // offset inside output i's window
          if (kk >= 0 && kk <= b) st[i].step(l, r, kk >= in.a, rp, in);
        }
      }
    }
    V o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t t = c.e0 + p0 + i;
      int64_t last = c.T - 1 - t;  // min(b, T-1-t): api.py:284-293 + classic.py:59
      if (!in.unbounded && last > in.b) last = in.b;
      const int ncand = (int)(last - in.a + 1);
      if (ncand <= 0 || p0 + i >= nlive) {
#pragma unroll
        for (int w = 0; w < VW; ++w) o[i].x[w] = f_nan();
      } else {
        o[i] = st[i].done(ncand, rp);
        if constexpr (SEM == SEM_SMOOTH) {
          // the shifted sums lost precision (huge dynamic range inside the tile): exact, slow
          if (!(st[i].acc >= kSmoothTiny) || !(st[i].acc <= 1.0e30f)) {
            // every term 1/(S_i + e^{-r_i/tau}) gets its own scale: (max, sum) pairs for the
            // running prefix S and for the accumulated candidates
            const float* pl = c.scr(0);
            const float* pr = c.scr(1);
            const int p = p0 + i;
            float mS = -f_inf(), S = 0.f, mA = -f_inf(), A = 0.f;
            for (int kk = 0; kk <= (int)last; ++kk) {
              const float zl = -pl[pidx(p + kk)] * rp.k;
              if (zl > mS) {
                S = S * fast_ex2(mS - zl) + 1.f;
                mS = zl;
              } else {
                S += fast_ex2(zl - mS);
              }
              if (kk >= in.a) {
                const float zr = -pr[pidx(p + kk)] * rp.k;
                const float mc = fmaxf(mS, zr);
                const float cl = -(mc + fast_lg2(S * fast_ex2(mS - mc) + fast_ex2(zr - mc)));  // log2 of the term
                if (cl > mA) {
                  A = A * fast_ex2(mA - cl) + 1.f;
                  mA = cl;
                } else {
                  A += fast_ex2(cl - mA);
                }
              }
            }
            o[i].x[0] = (mA + fast_lg2(A)) * rp.invk;
          }
        }
      }
    }
    store_val8<SEM>(c, in.dst, tid, o);
  }
  ex.sync();
}

// ------------------------------------------------------------------ instruction dispatch
template <int SEM, class EX>
PYSTL_HD void run_instr(EX& ex, const Ctx& c, const TInstr& in) {
  constexpr bool has_sum = (SEM == SEM_CUMULATIVE);
  switch (in.op) {
    case T_NARY:
      if (in.mode == MODE_AND) op_nary<SEM, MODE_AND>(ex, c, in);
      else op_nary<SEM, MODE_OR>(ex, c, in);
      break;
    case T_WIN_DIRECT:
      if (in.mode == MODE_AND) op_win_direct<SEM, MODE_AND>(ex, c, in);
      else if (in.mode == MODE_OR) op_win_direct<SEM, MODE_OR>(ex, c, in);
      else if constexpr (has_sum) op_win_direct<SEM, MODE_SUM>(ex, c, in);
      break;
    case T_WIN_VH:
      if (in.mode == MODE_AND) op_win_vh<SEM, MODE_AND>(ex, c, in);
      else if (in.mode == MODE_OR) op_win_vh<SEM, MODE_OR>(ex, c, in);
      else if constexpr (has_sum) op_win_vh<SEM, MODE_SUM>(ex, c, in);
      break;
    case T_WIN_SUFFIX:
      if (in.mode == MODE_AND) op_win_suffix<SEM, MODE_AND>(ex, c, in);
      else if (in.mode == MODE_OR) op_win_suffix<SEM, MODE_OR>(ex, c, in);
      else if constexpr (has_sum) op_win_suffix<SEM, MODE_SUM>(ex, c, in);
      break;
    case T_UNTIL:
      op_until<SEM>(ex, c, in);
      break;
    case T_CLAMP_SCAN:
      if constexpr (SEM == SEM_CLASSIC) op_clamp_scan(ex, c, in);
      break;
    default:
      break;
  }
}

// Evaluate the whole program on a tile whose raw samples are already in shared memory.
template <int SEM, class EX>
PYSTL_HD void run_program(EX& ex, const Ctx& c, const float* raw) {
  eval_predicates<SEM>(ex, c, raw);
  ex.sync();
  const TileProgram& P = *c.P;
  for (int i = 0; i < P.n_instr; ++i) run_instr<SEM>(ex, c, P.ins[i]);
}

// ------------------------------------------------------------------ result
// Write the formula value of this tile's outputs [0, n_out) to out_trace[(e0 + p) * C + w].
// Outputs the reference raises on (api.py:218-221) are NaN.  Optionally accumulates the
// sum of the valid outputs (all channels) for the fused mean-robustness reduction.
template <int SEM, class EX>
PYSTL_HD void store_root(EX& ex, const Ctx& c, float* out_trace, int n_out, double* mean_acc) {
  constexpr int VW = Tr<SEM>::VW;
  const TileProgram& P = *c.P;
  const int64_t last_valid = P.raises_always ? -1 : c.T - 1 - P.start_slack;
  PYSTL_FOR_THREADS(ex, tid) {
    Vec<VW> v[8];
    load_val8<SEM>(c, P.root, P.root_neg != 0, tid, v);
    float part = 0.f;
    const int p0 = 8 * tid;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t t = c.e0 + p0 + i;
      if (p0 + i < n_out && t < c.T) {
        const bool ok = t <= last_valid;
#pragma unroll
        for (int w = 0; w < VW; ++w) {
          const float x = ok ? v[i].x[w] : f_nan();
          out_trace[t * VW + w] = x;
          if (ok) part += x;
        }
      }
    }
    ex.put(tid, part);
  }
  if (mean_acc) ex.add_block_sum(c.red(), mean_acc);
}

// ------------------------------------------------------------------ host executor (emulation)
#if !defined(__CUDACC__)
struct HostExec {
  int nt_;
  float stash[kMaxThreads];
  explicit HostExec(int nt) : nt_(nt) {}
  int t0() const { return 0; }
  int t1() const { return nt_; }
  void sync() {}
  void put(int tid, float v) { stash[tid] = v; }
  float all_max(float*) {
    float m = -INFINITY;
    for (int i = 0; i < nt_; ++i) m = fmaxf(m, stash[i]);
    return m;
  }
  void add_block_sum(float*, double* acc) {
    double s = 0;
    for (int i = 0; i < nt_; ++i) s += stash[i];
    *acc += s;
  }
  // a[w*n + c] <- comb(a[c+1], ..., a[n-1]) (identity for the last chunk), in time order
  template <int AW, class OP>
  void excl_suffix_scan(float* a, int n, float*) {
    typename OP::A run = OP::ident();
    for (int ch = n - 1; ch >= 0; --ch) {
      typename OP::A cur;
      for (int w = 0; w < AW; ++w) cur.x[w] = a[w * n + ch];
      for (int w = 0; w < AW; ++w) a[w * n + ch] = run.x[w];
      run = OP::comb(cur, run);
    }
  }
};
#endif

}  // namespace pystl
