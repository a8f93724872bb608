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
#if defined(__CUDACC_RTC__)
// run-time compilation (jit.cpp): no host headers
#define INFINITY __int_as_float(0x7f800000)
#define NAN __int_as_float(0x7fffffff)
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#endif

#include "program.h"

#if defined(__CUDACC__)
#define PYSTL_HD __host__ __device__ __forceinline__
#else
#define PYSTL_HD inline
#endif

// D-GMSR exponent p: a run-time parameter of the interpreter tier; the fused tier bakes it into the
// generated kernel (PYSTL_CONST_P) so that the p == 1 fast paths fold at compile time.
#if defined(PYSTL_CONST_P)
#define PYSTL_P(rp) (PYSTL_CONST_P)
#else
#define PYSTL_P(rp) ((rp).p)
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

// bit casts (values of SEM_CLASSIC_SRC carry an integer in a float register)
PYSTL_HD int f2i(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  int i;
  memcpy(&i, &f, 4);
  return i;
#endif
}
PYSTL_HD float i2f(int i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  float f;
  memcpy(&f, &i, 4);
  return f;
#endif
}

// Host builds only (tests/emu): PYSTL_EMU_MUFU=<mask> perturbs the libm results by the documented
// error bounds of the hardware approximations (1: ex2 2^-22 relative, 2: lg2 2^-22 absolute inside
// (0.5, 2) and relative outside, 4: rcp 2^-23, 8: sqrt 2^-23, 16: flush denormals), with a sign and
// size derived from the argument's bits, so that the CPU tier can measure how sensitive a formula
// is to them (tests/error_table.py --backend emu).
#if !defined(__CUDA_ARCH__) && defined(PYSTL_EMU_MUFU)
inline float emu_unit(float x) {  // deterministic pseudo-random number in [-1, 1] from the bits of x
  unsigned u;
  memcpy(&u, &x, 4);
  u = (u ^ 0x9e3779b9u) * 0x85ebca6bu;
  u ^= u >> 13;
  u *= 0xc2b2ae35u;
  u ^= u >> 16;
  return (float)((double)u / 2147483648.0 - 1.0);
}
inline float emu_ftz(float x) { return ((PYSTL_EMU_MUFU & 16) && fabsf(x) < 1.17549435e-38f) ? 0.f * x : x; }
#define PYSTL_EMU_REL(bit, y, x, mag) (((PYSTL_EMU_MUFU) & (bit)) ? (y) * (1.f + (mag) * emu_unit(x)) : (y))
#else
#define PYSTL_EMU_REL(bit, y, x, mag) (y)
inline float emu_ftz(float x) { return x; }
#endif

// Device builds: PYSTL_PRECISE_MATH=<mask> (an NVRTC flag, experiments only) swaps an approximation
// for the correctly rounded / 1-2 ulp library function (1: ex2, 2: lg2, 4: rcp, 8: sqrt).
#if !defined(PYSTL_PRECISE_MATH)
#define PYSTL_PRECISE_MATH 0
#endif
PYSTL_HD float fast_ex2(float x) {
#if defined(__CUDA_ARCH__)
  if (PYSTL_PRECISE_MATH & 1) return exp2f(x);
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return emu_ftz(PYSTL_EMU_REL(1, exp2f(emu_ftz(x)), x, 2.3841858e-7f));
#endif
}
PYSTL_HD float fast_lg2(float x) {
#if defined(__CUDA_ARCH__)
  if (PYSTL_PRECISE_MATH & 2) return log2f(x);
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
#if defined(PYSTL_EMU_MUFU)
  if ((PYSTL_EMU_MUFU & 2) && x > 0.5f && x < 2.f) return log2f(x) + 2.3841858e-7f * emu_unit(x);
#endif
  return PYSTL_EMU_REL(2, log2f(emu_ftz(x)), x, 2.3841858e-7f);
#endif
}

// Hardware transcendentals (MUFU): 2-ulp ex2 / lg2 / rcp / sqrt.  The parity bar of the smooth,
// agm and dgmsr semantics is |err| <= 5e-6 + 1e-5 |ref| (tests/parity_util.py); the absolute
// error of lg2.approx is below 2^-22 and window means do not amplify it.
PYSTL_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
  if (PYSTL_PRECISE_MATH & 4) return __frcp_rn(x);
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return emu_ftz(PYSTL_EMU_REL(4, 1.f / emu_ftz(x), x, 1.1920929e-7f));
#endif
}
PYSTL_HD float fast_sqrt(float x) {
#if defined(__CUDA_ARCH__)
  if (PYSTL_PRECISE_MATH & 8) return __fsqrt_rn(x);
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return PYSTL_EMU_REL(8, sqrtf(emu_ftz(x)), x, 1.1920929e-7f);
#endif
}
constexpr float kLn2 = 0.69314718055994531f;
constexpr float kLog2e = 1.4426950408889634f;
PYSTL_HD float fast_ln(float x) { return fast_lg2(x) * kLn2; }
PYSTL_HD float fast_exp(float x) { return fast_ex2(x * kLog2e); }
PYSTL_HD float fast_log1p(float x) { return fast_lg2(1.f + x) * kLn2; }
PYSTL_HD float fast_expm1(float x) { return fast_ex2(x * kLog2e) - 1.f; }
PYSTL_HD float fast_pow(float x, float y) { return fast_ex2(y * fast_lg2(x)); }  // x > 0

// ------------------------------------------------------------------ values and accumulators
template <int SEM>
struct Tr {
  static constexpr int VW = (SEM == SEM_CUMULATIVE || SEM == SEM_CLASSIC_SRC) ? 2 : 1;  // words per node value
  static constexpr int AW = (SEM == SEM_CLASSIC || SEM == SEM_SMOOTH) ? 1 : 2;          // words per accumulator
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
  } else if constexpr (SEM == SEM_CLASSIC_SRC) {
    r.x[0] = -v.x[0];
    r.x[1] = i2f(f2i(v.x[1]) ^ (int)0x80000000);  // the source's sign bit
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

// Classic with provenance: x[0] = value, x[1] = bits of (sign << 31 | predicate << 24 | tile
// position).  Ties go to the earlier operand, like torch.min / max(dim) (torch/semantics.py:106,113).
template <int MODE>
struct Red<SEM_CLASSIC_SRC, MODE> {
  using V = Vec<2>;
  using A = Vec<2>;
  static PYSTL_HD A ident() { return A{{MODE == MODE_AND ? f_inf() : -f_inf(), 0.f}}; }
  static PYSTL_HD A lift(V v, const RP&, float) { return v; }
  static PYSTL_HD A comb(A a, A b) {
    const bool take_b = MODE == MODE_AND ? (b.x[0] < a.x[0]) : (b.x[0] > a.x[0]);
    return A{{take_b ? b.x[0] : a.x[0], take_b ? b.x[1] : a.x[1]}};
  }
  static PYSTL_HD V fin(A a, float, const RP&) { return a; }
};

template <int MODE>
struct Red<SEM_SMOOTH, MODE> {
  using V = Vec<1>;
  using A = Vec<1>;
  static PYSTL_HD float expo(V v, const RP& rp) { return (MODE == MODE_AND ? -v.x[0] : v.x[0]) * rp.k; }
  static PYSTL_HD A ident() { return A{{0.f}}; }
  // one FFMA + one MUFU each way: 2^(+-k v - shift), and +-invk (lg2 S + shift)
  static PYSTL_HD A lift(V v, const RP& rp, float) { return A{{fast_ex2(fmaf(v.x[0], MODE == MODE_AND ? -rp.k : rp.k, -rp.shift))}}; }
  static PYSTL_HD A comb(A a, A b) { return A{{a.x[0] + b.x[0]}}; }
  static PYSTL_HD V fin(A a, float, const RP& rp) {
    const float sk = MODE == MODE_AND ? -rp.invk : rp.invk;
    return V{{fmaf(fast_lg2(a.x[0]), sk, rp.shift * sk)}};
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
      return (x <= 0.f) ? A{{w * x, f_inf()}} : A{{0.f, w * fast_log1p(x)}};
    } else {  // agm.py:86-88
      return (x > 0.f) ? A{{w * x, f_inf()}} : A{{0.f, w * fast_log1p(-x)}};
    }
  }
  static PYSTL_HD A comb(A a, A b) { return A{{a.x[0] + b.x[0], a.x[1] + b.x[1]}}; }
  static PYSTL_HD V fin(A a, float sw, const RP&) {
    const float inv = fast_rcp(sw);
    const float g = fast_expm1(a.x[1] * inv);  // (+inf when an entry took the arithmetic branch: discarded)
    return V{{(a.x[1] == f_inf()) ? a.x[0] * inv : (MODE == MODE_AND ? g : -g)}};
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
    const float x2 = x * x;
    const bool neg = x <= 0.f;
    const float pw = PYSTL_P(rp) == 1 ? x2 : (x2 > 0.f ? fast_pow(x2, (float)PYSTL_P(rp)) : 0.f);
    // (the logarithm of a non-positive x is discarded by the select)
    return A{{neg ? w * pw : 0.f, neg ? f_inf() : 2.f * w * fast_ln(x)}};
  }
  static PYSTL_HD A comb(A a, A b) { return A{{a.x[0] + b.x[0], a.x[1] + b.x[1]}}; }
  static PYSTL_HD V fin(A a, float sw, const RP& rp) {
    const float inv = fast_rcp(sw);
    // _dgmsr.py:83-87
    const float m = rp.eps_p + a.x[0] * inv;
    const float hn = rp.sqrt_eps - (PYSTL_P(rp) == 1 ? fast_sqrt(m) : fast_pow(m, rp.inv2p));
    // _dgmsr.py:96-102, in the log domain: ln m0 = logaddexp(sw ln eps, sum 2w ln v) / sw
    const float u = sw * rp.ln_eps, s = a.x[1];
    const float lm0 = (fmaxf(u, s) + fast_log1p(fast_exp(-fabsf(u - s)))) * inv;
    const float hp = fast_exp(0.5f * lm0) - rp.sqrt_eps;
    const float h = (a.x[1] == f_inf()) ? hn : hp;
    return V{{MODE == MODE_AND ? h : -h}};
  }
};

// ------------------------------------------------------------------ tile context
struct Ctx {
  float* smem;
  int nt, lext, D;
  int off_planes, off_scratch, off_chunk, off_red;
  int gl, gr;   // scratch guard bands (floats), see TileGeom
  int n_slots;  // value slots (adjoint planes of the backward kernel follow them)
  int nlive;    // tile samples that exist in the trace: min(T - e0, lext)
  int first;    // tile position of trace sample 0 (= -e0, clamped): positions below it do not exist
  int last_ok;  // last tile position whose output does not raise (T-1-start_slack-e0, clamped), < 0: none
  RP rp;
  PYSTL_HD float* plane(int slot, int w, int vw) const { return smem + off_planes + (slot * vw + w) * lext; }
  // scratch plane i / chunk-aggregate array i, addressed by LOGICAL position: valid for
  // positions [-gl, lext + gr) and chunks [-gl/8, nt + gr/8)
  PYSTL_HD float* scr(int i) const { return smem + off_scratch + i * (gl + lext + gr) + gl; }
  PYSTL_HD float* chunk(int i) const { return smem + off_chunk + i * (gl / 8 + nt + gr / 8) + gl / 8; }
  PYSTL_HD float* raw() const { return smem + off_scratch; }
  // [lo, hi] of every value plane (smooth only): every thread writes the same numbers, so a
  // thread always reads back what it wrote itself and no barrier is needed
  PYSTL_HD float* bounds() const { return smem + off_red + 128; }
  // backward: adj_tab()[s] = plane index of the adjoint of value slot s (copied from the program)
  PYSTL_HD int* adj_tab() const { return reinterpret_cast<int*>(smem + off_red + 128 + 2 * (n_slots + 1)); }
  PYSTL_HD float* red() const { return smem + off_red; }
};

PYSTL_HD Ctx make_ctx(float* smem, const TileGeom& g, const TileProgram& P, int64_t e0) {
  Ctx c;
  c.smem = smem;
  c.nt = g.nt;
  c.lext = g.lext;
  c.D = g.D;
  c.off_planes = g.off_planes;
  c.off_scratch = g.off_scratch;
  c.off_chunk = g.off_chunk;
  c.off_red = g.off_red;
  c.gl = g.gl;
  c.gr = g.gr;
  c.n_slots = g.n_slots;
  const int64_t live = g.T - e0;
  c.nlive = live < (int64_t)g.lext ? (int)live : g.lext;
  c.first = e0 > (int64_t)(1 << 30) ? -(1 << 30) : -(int)e0;
  const int64_t ok = P.raises_always ? -1 : g.T - 1 - P.start_slack - e0;
  c.last_ok = ok > (int64_t)(1 << 30) ? (1 << 30) : (ok < -1 ? -1 : (int)ok);
  c.rp = P.rp;
  return c;
}

// smooth only: an upper bound of the operand's exponent over the tile, from the value range
// tracked per plane (exact for predicates, within log2(#terms) per operator above them)
template <int MODE>
PYSTL_HD float smooth_shift(const Ctx& c, int slot, bool neg) {
  const float lo = c.bounds()[2 * slot], hi = c.bounds()[2 * slot + 1];
  const float vlo = neg ? -hi : lo, vhi = neg ? -lo : hi;  // range of the (possibly negated) operand
  return (MODE == MODE_AND ? -vlo : vhi) * c.rp.k;
}
// range of softmin / softmax over n terms of operands within [lo, hi]:
//   min - tau ln n <= softmin <= min,   max <= softmax <= max + tau ln n
template <int MODE>
PYSTL_HD void smooth_set_bounds(const Ctx& c, int dst, float lo, float hi, float n) {
  const float pad = c.rp.invk * fast_lg2(n > 1.f ? n : 1.f) + 1e-6f * (fabsf(lo) + fabsf(hi));
  c.bounds()[2 * dst] = MODE == MODE_AND ? lo - pad : lo;
  c.bounds()[2 * dst + 1] = MODE == MODE_AND ? hi : hi + pad;
}

// own chunk of a plane <-> registers
PYSTL_HD void load8(const float* plane, int tid, float (&v)[8]) {  // tid = chunk index, may be negative
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

// run of 8 consecutive logical samples starting at `start` (any alignment, any sign) from a
// scratch plane.  No bounds checks: the guard bands of the scratch planes hold the identity.
PYSTL_HD void load_run8(const float* plane, int start, float (&v)[8]) {
  const F4* P4 = reinterpret_cast<const F4*>(plane);
  const int f0 = start >> 2;  // arithmetic shift: floor
  const int r = start & 3;
  float t[12];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const F4 q = P4[pf4(f0 + j)];
#pragma unroll
    for (int k = 0; k < 4; ++k) t[4 * j + k] = q.v[k];
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
// Strided ownership (sample p = tid + nt*i): consecutive lanes read consecutive rows of the
// raw [sample][channel] tile (one vector load per row for D = 1, 2, 4, 8) and write
// consecutive words of the predicate plane.  Branch-free: rows past the horizon are
// evaluated on whatever the staging area holds and never consumed.
template <int DV>
PYSTL_HD void load_row(const float* raw, int p, int D, float (&row)[8]) {
  if constexpr (DV == 1) {
    row[0] = raw[p];
  } else if constexpr (DV == 2) {
    struct alignas(8) F2 { float v[2]; };
    const F2 q = reinterpret_cast<const F2*>(raw)[p];
    row[0] = q.v[0];
    row[1] = q.v[1];
  } else if constexpr (DV == 4) {
    const F4 q = reinterpret_cast<const F4*>(raw)[p];
#pragma unroll
    for (int k = 0; k < 4; ++k) row[k] = q.v[k];
  } else if constexpr (DV == 8) {
    const F4 q0 = reinterpret_cast<const F4*>(raw)[2 * p], q1 = reinterpret_cast<const F4*>(raw)[2 * p + 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      row[k] = q0.v[k];
      row[4 + k] = q1.v[k];
    }
  }
  (void)D;
}
// row[ch] with ch uniform: a select chain keeps the row in registers
template <int DV>
PYSTL_HD float pick(const float (&row)[8], int ch) {
  float x = row[0];
#pragma unroll
  for (int d = 1; d < DV; ++d) x = (ch == d) ? row[d] : x;
  return x;
}

template <int DV>
PYSTL_HD float chan(const float (&row)[8], const float* raw, int p, int D, int ch) {
  if constexpr (DV > 0) return pick<DV>(row, ch);
  else return raw[p * D + ch];
}

template <int SEM, int DV, class EX>
PYSTL_HD void eval_predicates_d(EX& ex, const Ctx& c, const TileProgram& P, const float* raw) {
  constexpr int VW = Tr<SEM>::VW;
  const int D = c.D, nt = c.nt, nlive = c.nlive;
  const int np = P.n_preds;
#pragma unroll 1
  for (int q = 0; q < np; ++q) {
    // everything about the predicate is fetched once, then 8 samples are streamed through it
    const int nterms = P.preds[q].n_terms;
    const float bias = P.preds[q].bias;
    const bool f64 = (SEM == SEM_CLASSIC || SEM == SEM_CLASSIC_SRC) && P.preds[q].f64;
    const int ch0 = P.preds[q].ch[0], ch1 = P.preds[q].ch[1];
    const float c0 = P.preds[q].coeff[0], c1 = nterms > 1 ? P.preds[q].coeff[1] : 0.f;
    float* pl0 = c.plane(P.preds[q].dst, 0, VW);
    float* pl1 = c.plane(P.preds[q].dst, VW - 1, VW);
    PYSTL_FOR_THREADS(ex, tid) {
      float lo = f_inf(), hi = -f_inf();
#pragma unroll
      for (int i = 0; i < kChunk; ++i) {
        const int p = tid + nt * i;
        float row[8];
        if constexpr (DV > 0) load_row<DV>(raw, p, D, row);
        float acc;
        if (!f64 && nterms <= 2) {
          // 0 + (+-x) is exact: a single-term predicate rounds once
          acc = fmaf(c0, chan<DV>(row, raw, p, D, ch0), 0.f);
          if (nterms == 2) acc = fmaf(c1, chan<DV>(row, raw, p, D, ch1), acc);
          if (nterms == 0) acc = 0.f;
          acc += bias;
        } else if (f64) {
          // the reference evaluates predicates in float64 (classic.py:20); same operations, same
          // order as formula.affine_predicate, one final rounding: bit-exact after the cast
          double a = 0.0;
#pragma unroll 1
          for (int k = 0; k < nterms; ++k)
            a = fma((double)P.preds[q].coeff[k], (double)chan<DV>(row, raw, p, D, P.preds[q].ch[k]), a);
          acc = (float)(a + (double)bias);
        } else {
          acc = 0.f;
#pragma unroll 1
          for (int k = 0; k < nterms; ++k)
            acc = fmaf(P.preds[q].coeff[k], chan<DV>(row, raw, p, D, P.preds[q].ch[k]), acc);
          acc += bias;
        }
        const int pp = pidx(p);
        if constexpr (SEM == SEM_CUMULATIVE) {  // cumulative.py:29-30
          pl0[pp] = fmaxf(0.f, acc);
          pl1[pp] = fminf(0.f, acc);
        } else if constexpr (SEM == SEM_CLASSIC_SRC) {
          pl0[pp] = acc;
          pl1[pp] = i2f((q << 24) | p);
        } else {
          pl0[pp] = acc;
        }
        if constexpr (SEM == SEM_SMOOTH) {
          if (p < nlive) {
            lo = fminf(lo, acc);
            hi = fmaxf(hi, acc);
          }
        }
      }
      if constexpr (SEM == SEM_SMOOTH) ex.put2(tid, -lo, hi);
    }
    if constexpr (SEM == SEM_SMOOTH) {
      // value range of the plane: smooth windows shift their exponentials by a bound of the max
      float mlo, mhi;
      ex.all_max2(c.red(), mlo, mhi);
      c.bounds()[2 * P.preds[q].dst] = -mlo;
      c.bounds()[2 * P.preds[q].dst + 1] = mhi;
    }
  }
}

template <int SEM, class EX>
PYSTL_HD void eval_predicates(EX& ex, const Ctx& c, const TileProgram& P, const float* raw) {
  switch (c.D) {
    case 1: eval_predicates_d<SEM, 1>(ex, c, P, raw); break;
    case 2: eval_predicates_d<SEM, 2>(ex, c, P, raw); break;
    case 4: eval_predicates_d<SEM, 4>(ex, c, P, raw); break;
    case 8: eval_predicates_d<SEM, 8>(ex, c, P, raw); break;
    default: eval_predicates_d<SEM, 0>(ex, c, P, raw); break;
  }
}

// ------------------------------------------------------------------ n-ary And / Or
template <int SEM, int MODE, class EX>
PYSTL_HD void op_nary(EX& ex, const Ctx& c, const TInstr& in) {
  using R = Red<SEM, MODE>;
  using V = typename R::V;
  using A = typename R::A;
  const int nsrc = in.nsrc, dst = in.dst;
  const unsigned negmask = in.negmask;
  const float sw = in.sw;
  if constexpr (SEM == SEM_SMOOTH) {
    float lo = MODE == MODE_AND ? f_inf() : -f_inf(), hi = lo;  // min / max over the operands' ranges
    for (int j = 0; j < nsrc; ++j) {
      const bool ng = (negmask >> j) & 1;
      const float l = ng ? -c.bounds()[2 * in.src[j] + 1] : c.bounds()[2 * in.src[j]];
      const float h = ng ? -c.bounds()[2 * in.src[j]] : c.bounds()[2 * in.src[j] + 1];
      lo = MODE == MODE_AND ? fminf(lo, l) : fmaxf(lo, l);
      hi = MODE == MODE_AND ? fminf(hi, h) : fmaxf(hi, h);
    }
    smooth_set_bounds<MODE>(c, dst, lo, hi, (float)nsrc);
  }
  PYSTL_FOR_THREADS(ex, tid) {
    RP rp = c.rp;
    A acc[8];
    if constexpr (SEM == SEM_SMOOTH) {
      // exact max-shifted logsumexp per sample (smooth.py:18-22): pass 1 finds the max exponent
      float m[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = -f_inf();
#pragma unroll 1
      for (int j = 0; j < nsrc; ++j) {
        V v[8];
        load_val8<SEM>(c, in.src[j], (negmask >> j) & 1, tid, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], R::expo(v[i], rp));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = R::ident();
#pragma unroll 1
      for (int j = 0; j < nsrc; ++j) {
        V v[8];
        load_val8<SEM>(c, in.src[j], (negmask >> j) & 1, tid, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i].x[0] += fast_ex2(R::expo(v[i], rp) - m[i]);
      }
      V o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        rp.shift = m[i];
        o[i] = R::fin(acc[i], sw, rp);
      }
      store_val8<SEM>(c, dst, tid, o);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = R::ident();
#pragma unroll 1
      for (int j = 0; j < nsrc; ++j) {
        V v[8];
        load_val8<SEM>(c, in.src[j], (negmask >> j) & 1, tid, v);
        const float wj = in.w[j];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = R::comb(acc[i], R::lift(v[i], rp, wj));
      }
      V o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = R::fin(acc[i], sw, rp);
      store_val8<SEM>(c, dst, tid, o);
    }
  }
}

// A window instruction, decoded once per thread into registers.
struct Win {
  int a, b;    // offsets relative to t
  int src;     // operand plane slot
  bool neg;    // operand negated
  int dst;
};
PYSTL_HD Win decode_win(const TInstr& in) { return Win{in.a, in.b, in.src[0], (in.negmask & 1) != 0, in.dst}; }

// number of samples of the clipped window [t+a, min(t+b, T-1)] (api.py:49-54), in tile positions
PYSTL_HD int window_count(const Ctx& c, int p, int a, int b) {
  int lo = p + a, hi = p + b;
  if (lo < c.first) lo = c.first;
  if (hi > c.nlive - 1) hi = c.nlive - 1;
  return hi - lo + 1;
}

// smooth only: exact logsumexp of one window straight from the operand plane (rare path,
// taken when the block-shifted sum lost its precision to underflow)
template <int SEM, int MODE>
PYSTL_HD Vec<1> smooth_exact_window(const Ctx& c, const Win& w, int p) {
  using R = Red<SEM, MODE>;
  const float* pl = c.plane(w.src, 0, 1);
  int lo = p + w.a, hi = p + w.b;
  if (lo < 0) lo = 0;
  if (hi > c.nlive - 1) hi = c.nlive - 1;
  float m = -f_inf();
  for (int q = lo; q <= hi; ++q) {
    Vec<1> v{{w.neg ? -pl[pidx(q)] : pl[pidx(q)]}};
    m = fmaxf(m, R::expo(v, c.rp));
  }
  float s = 0.f;
  for (int q = lo; q <= hi; ++q) {
    Vec<1> v{{w.neg ? -pl[pidx(q)] : pl[pidx(q)]}};
    s += fast_ex2(R::expo(v, c.rp) - m);
  }
  RP rp = c.rp;
  rp.shift = m;
  return R::fin(Vec<1>{{s}}, 1.f, rp);
}

constexpr float kSmoothTiny = 7.8886090522101181e-31f;  // 2^-100

// Per-thread view of the window sizes of its 8 outputs: away from the ends of the trace every
// window has b-a+1 samples (`interior`), so the clipped count is only computed near the horizon.
struct WinCount {
  bool interior;
  float full;
};
PYSTL_HD WinCount win_count8(const Ctx& c, const Win& w, int p0) {
  WinCount k;
  k.interior = (p0 + 7 + w.b <= c.nlive - 1) && (p0 + w.a >= c.first);
  k.full = (float)(w.b - w.a + 1);
  return k;
}

// accumulators of the thread's 8 outputs -> values.  Thread-level fast path for windows that
// are not clipped; the clipped count, the NaN of an empty window (the reference raises,
// api.py:218-221) and smooth's exact recomputation are handled off the hot path.
template <int SEM, int MODE>
PYSTL_HD void finish8(const Ctx& c, const Win& w, const RP& rp, const typename Red<SEM, MODE>::A (&acc)[8], int p0,
                      const WinCount& k, typename Red<SEM, MODE>::V (&o)[8]) {
  using R = Red<SEM, MODE>;
  if (k.interior) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = R::fin(acc[i], k.full, rp);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = window_count(c, p0 + i, w.a, w.b);
      o[i] = R::fin(acc[i], (float)(n > 0 ? n : 1), rp);
      if (n <= 0) {
#pragma unroll
        for (int q = 0; q < Tr<SEM>::VW; ++q) o[i].x[q] = f_nan();
      }
    }
  }
  if constexpr (SEM == SEM_SMOOTH) {
    bool lost = false;
#pragma unroll
    for (int i = 0; i < 8; ++i) lost = lost || !(acc[i].x[0] >= kSmoothTiny);
    if (lost) {  // rare: the block-shifted sum underflowed (or the window is empty)
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        float a_i = 0.f;  // acc[i] with a static register index
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j == i) a_i = acc[j].x[0];
        if (!(a_i >= kSmoothTiny)) {
          const float v = window_count(c, p0 + i, w.a, w.b) <= 0 ? f_nan() : smooth_exact_window<SEM, MODE>(c, w, p0 + i).x[0];
          // (static indexing keeps o[] in registers)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j == i) o[j].x[0] = v;
        }
      }
    }
  }
}

// lift the own chunk of the operand, treating samples past the horizon as the identity
template <int SEM, int MODE>
PYSTL_HD void lift_own8(const Ctx& c, const Win& w, const RP& rp, int tid, typename Red<SEM, MODE>::A (&x)[8]) {
  using R = Red<SEM, MODE>;
  typename R::V v[8];
  load_val8<SEM>(c, w.src, w.neg, tid, v);
  const int nlive = c.nlive;
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (8 * tid + i < nlive) ? R::lift(v[i], rp, 1.f) : R::ident();
}

// Write the identity into the guard bands of `nplanes` scratch planes (and, if `chunks`, of the
// chunk-aggregate arrays).  A few threads do it; it is part of the instruction's first phase.
template <int AW>
PYSTL_HD void fill_guards(const Ctx& c, int tid, Vec<AW> id, int nplanes_per_word, bool chunks) {
  const int ncl = c.gl >> 3, ncr = c.gr >> 3;
  for (int g = tid; g < ncl + ncr; g += c.nt) {
    const int ch = g < ncr ? c.nt + g : -(g - ncr) - 1;
#pragma unroll
    for (int w = 0; w < AW; ++w) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = id.x[w];
      for (int k = 0; k < nplanes_per_word; ++k) store8(c.scr(k * AW + w), ch, t);
      if (chunks) c.chunk(w)[ch] = id.x[w];
    }
  }
}

// clamp a window offset to the tile: beyond +-lext every position is past the horizon anyway
PYSTL_HD int clamp_off(int x, int lext) { return x > lext ? lext : (x < -lext ? -lext : x); }

// ------------------------------------------------------------------ sliding windows
// aggregate of chunk `ch` (the guard entries hold the identity)
template <int AW>
PYSTL_HD Vec<AW> load_agg(const Ctx& c, int ch) {
  Vec<AW> r;
#pragma unroll
  for (int w = 0; w < AW; ++w) r.x[w] = c.chunk(w)[ch];
  return r;
}

// The two phases of a sliding-window reduction out[p] = (+)_{k in [wa, wb]} x[p + k] over the
// tile, for any associative (+) (R::comb, identity R::ident; operands are always combined in
// time order).  A __syncthreads separates them.
//
//   direct   (wb - wa + 1 <= 8)  phase 1 stages x; phase 2 walks the 8 + W - 1 samples the
//            thread's outputs need as float4s and feeds every output it belongs to
//   van Herk / Gil-Werman (>= 9) chunk = the 8 samples of one thread.  Phase 1: in-register
//            suffix and prefix scans of the own chunk -> scratch planes SUF, PRE and the chunk
//            aggregate AGG.  Phase 2: out[p] = SUF[p+wa] (+) AGG[whole chunks strictly inside]
//            (+) PRE[p+wb]; the pieces are disjoint, so (+) need not be idempotent.
template <class R, int AW>
PYSTL_HD void win_phase1(const Ctx& c, int tid, bool vh, const Vec<AW> (&x)[8]) {
  if (!vh) {
#pragma unroll
    for (int w = 0; w < AW; ++w) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = x[i].x[w];
      store8(c.scr(w), tid, t);
    }
    fill_guards<AW>(c, tid, R::ident(), 1, false);
    return;
  }
  Vec<AW> suf[8], pre[8];
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
  fill_guards<AW>(c, tid, R::ident(), 2, true);
}

template <class R, int AW>
PYSTL_HD void win_phase2(const Ctx& c, int tid, bool vh, int wa, int wb, Vec<AW> (&acc)[8]) {
  using A = Vec<AW>;
  const A id = R::ident();
  const int p0 = 8 * tid;
  if (!vh) {
    const int W = wb - wa + 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = id;
    const int f_lo = (p0 + wa) >> 2, f_hi = (p0 + 7 + wb) >> 2;
#pragma unroll 1
    for (int f = f_lo; f <= f_hi; ++f) {
      F4 q[AW];
#pragma unroll
      for (int w = 0; w < AW; ++w) q[w] = reinterpret_cast<const F4*>(c.scr(w))[pf4(f)];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        A e;
#pragma unroll
        for (int w = 0; w < AW; ++w) e.x[w] = q[w].v[k];
        const int d = 4 * f + k - p0 - wa;  // window slot of output 0 this element fills
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (d - i >= 0 && d - i < W) acc[i] = R::comb(acc[i], e);
      }
    }
    return;
  }
  const int qa = wa >> 3, ra = wa & 7, qb = wb >> 3, rb = wb & 7;
  float sufv[AW][8], prev[AW][8];
#pragma unroll
  for (int w = 0; w < AW; ++w) {
    load_run8(c.scr(w), p0 + wa, sufv[w]);
    load_run8(c.scr(AW + w), p0 + wb, prev[w]);
  }
  const A al = load_agg<AW>(c, tid + qa + 1), ar = load_agg<AW>(c, tid + qb);
  A core = id;
#pragma unroll 2
  for (int ch = tid + qa + 2; ch <= tid + qb - 1; ++ch) core = R::comb(core, load_agg<AW>(c, ch));
  const bool two = qb >= qa + 2;
  // the four possible sets of whole chunks between the two partial ones, in time order
  const A m_l = R::comb(al, core), m_r = R::comb(core, ar), m_lr = R::comb(m_l, ar);
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
    A mid;
#pragma unroll
    for (int w = 0; w < AW; ++w) mid.x[w] = inc_l ? (inc_r ? m_lr.x[w] : m_l.x[w]) : (inc_r ? m_r.x[w] : core.x[w]);
    acc[i] = R::comb(R::comb(s, mid), e);
  }
}

// Always / Eventually over a bounded interval (T_WIN_DIRECT, T_WIN_VH)
template <int SEM, int MODE, class EX>
PYSTL_HD void op_win(EX& ex, const Ctx& c, const TInstr& in, bool vh) {
  using R = Red<SEM, MODE>;
  using A = typename R::A;
  using V = typename R::V;
  constexpr int AW = Tr<SEM>::AW;
  RP rp = c.rp;
  const Win wn = decode_win(in);
  const int wa = clamp_off(wn.a, c.lext), wb = clamp_off(wn.b, c.lext);
  if constexpr (SEM == SEM_SMOOTH) {
    rp.shift = smooth_shift<MODE>(c, wn.src, wn.neg);
    const float lo = c.bounds()[2 * wn.src], hi = c.bounds()[2 * wn.src + 1];
    smooth_set_bounds<MODE>(c, wn.dst, wn.neg ? -hi : lo, wn.neg ? -lo : hi, (float)(wb - wa + 1));
  }
  PYSTL_FOR_THREADS(ex, tid) {
    A x[8];
    lift_own8<SEM, MODE>(c, wn, rp, tid, x);
    win_phase1<R, AW>(c, tid, vh, x);
  }
  ex.sync();
  PYSTL_FOR_THREADS(ex, tid) {
    A acc[8];
    win_phase2<R, AW>(c, tid, vh, wa, wb, acc);
    V o[8];
    finish8<SEM, MODE>(c, wn, rp, acc, 8 * tid, win_count8(c, wn, 8 * tid), o);
    store_val8<SEM>(c, wn.dst, tid, o);
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
  Win wn = decode_win(in);
  wn.b = 0x3fffffff;  // open end: clipped at the horizon by window_count
  if constexpr (SEM == SEM_SMOOTH) {
    rp.shift = smooth_shift<MODE>(c, wn.src, wn.neg);
    const float lo = c.bounds()[2 * wn.src], hi = c.bounds()[2 * wn.src + 1];
    smooth_set_bounds<MODE>(c, wn.dst, wn.neg ? -hi : lo, wn.neg ? -lo : hi, (float)c.lext);
  }
  PYSTL_FOR_THREADS(ex, tid) {
    A x[8];
    lift_own8<SEM, MODE>(c, wn, rp, tid, x);
    A agg = x[7];
#pragma unroll
    for (int i = 6; i >= 0; --i) agg = R::comb(x[i], agg);
#pragma unroll
    for (int w = 0; w < AW; ++w) c.chunk(w)[tid] = agg.x[w];
  }
  ex.sync();
  ex.template excl_suffix_scan<AW, RedOp<SEM, MODE>>(c.chunk(0), c.nt, (int)(c.chunk(1) - c.chunk(0)), c.red());
  PYSTL_FOR_THREADS(ex, tid) {
    A x[8], suf[8];
    lift_own8<SEM, MODE>(c, wn, rp, tid, x);
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
    fill_guards<AW>(c, tid, R::ident(), 1, false);
  }
  ex.sync();
  PYSTL_FOR_THREADS(ex, tid) {
    float sv[AW][8];
#pragma unroll
    for (int w = 0; w < AW; ++w) load_run8(c.scr(w), 8 * tid + clamp_off(wn.a, c.lext), sv[w]);
    V o[8];
    A acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int w = 0; w < AW; ++w) acc[i].x[w] = sv[w][i];
    }
    finish8<SEM, MODE>(c, wn, rp, acc, 8 * tid, WinCount{false, 0.f}, o);
    store_val8<SEM>(c, wn.dst, tid, o);
  }
  ex.sync();
}

// ------------------------------------------------------------------ classic unbounded until (clamp scan)
// W(s) = min(l_s, max(r_s, W(s+1))) is the composition of f_s(x) = min(A, max(B, x)),
// which is closed under composition, so the whole tile is one
// suffix scan.  Starting the recurrence at the tile's last sample is exact for every
// output whose bounded window lies inside the tile ("Classic until").
template <int SEM>
struct ClampOp;

template <>
struct ClampOp<SEM_CLASSIC> {
  using A = Vec<2>;  // (A, B)
  static constexpr int W = 2;
  static PYSTL_HD A ident() { return A{{f_inf(), -f_inf()}}; }
  static PYSTL_HD A comb(A f, A g) {  // f after g:  x -> f(g(x)), f is the EARLIER sample
    return A{{fminf(f.x[0], fmaxf(f.x[1], g.x[0])), fmaxf(f.x[1], g.x[1])}};
  }
  static PYSTL_HD A make(Vec<1> l, Vec<1> r) { return A{{l.x[0], r.x[0]}}; }
  static PYSTL_HD Vec<1> value(A f) { return Vec<1>{{fminf(f.x[0], f.x[1])}}; }  // f(-inf) = min(A, B)
};

template <>
struct ClampOp<SEM_CLASSIC_SRC> {
  using A = Vec<4>;  // (A, src A, B, src B); ties go to the earlier sample
  static constexpr int W = 4;
  static PYSTL_HD A ident() { return A{{f_inf(), 0.f, -f_inf(), 0.f}}; }
  static PYSTL_HD A comb(A f, A g) {
    const bool up = g.x[0] > f.x[2];  // max(B_f, A_g)
    const float mv = up ? g.x[0] : f.x[2], ms = up ? g.x[1] : f.x[3];
    const bool dn = mv < f.x[0];  // min(A_f, .)
    const bool bb = g.x[2] > f.x[2];
    return A{{dn ? mv : f.x[0], dn ? ms : f.x[1], bb ? g.x[2] : f.x[2], bb ? g.x[3] : f.x[3]}};
  }
  static PYSTL_HD A make(Vec<2> l, Vec<2> r) { return A{{l.x[0], l.x[1], r.x[0], r.x[1]}}; }
  static PYSTL_HD Vec<2> value(A f) {
    const bool b = f.x[2] < f.x[0];
    return Vec<2>{{b ? f.x[2] : f.x[0], b ? f.x[3] : f.x[1]}};
  }
};

struct Pair {  // a two-operand instruction, decoded once per thread
  int sl, sr, dst;
  bool nl, nr;
};
PYSTL_HD Pair decode_pair(const TInstr& in) {
  return Pair{in.src[0], in.src[1], in.dst, (in.negmask & 1) != 0, (in.negmask & 2) != 0};
}

template <int SEM>
PYSTL_HD void clamp_load_lr(const Ctx& c, const Pair& in, int tid, int nlive, typename ClampOp<SEM>::A (&x)[8]) {
  Vec<Tr<SEM>::VW> l[8], r[8];
  load_val8<SEM>(c, in.sl, in.nl, tid, l);
  load_val8<SEM>(c, in.sr, in.nr, tid, r);
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (8 * tid + i < nlive) ? ClampOp<SEM>::make(l[i], r[i]) : ClampOp<SEM>::ident();
}

template <int SEM, class EX>
PYSTL_HD void op_clamp_scan(EX& ex, const Ctx& c, const TInstr& ins) {
  using OP = ClampOp<SEM>;
  using A = typename OP::A;
  constexpr int W = OP::W;
  const int nlive = c.nlive;
  const Pair in = decode_pair(ins);
  PYSTL_FOR_THREADS(ex, tid) {
    A x[8];
    clamp_load_lr<SEM>(c, in, tid, nlive, x);
    A agg = x[7];
#pragma unroll
    for (int i = 6; i >= 0; --i) agg = OP::comb(x[i], agg);
#pragma unroll
    for (int w = 0; w < W; ++w) c.chunk(w)[tid] = agg.x[w];
  }
  ex.sync();
  ex.template excl_suffix_scan<W, OP>(c.chunk(0), c.nt, (int)(c.chunk(1) - c.chunk(0)), c.red());
  PYSTL_FOR_THREADS(ex, tid) {
    A x[8];
    clamp_load_lr<SEM>(c, in, tid, nlive, x);
    A run;
#pragma unroll
    for (int w = 0; w < W; ++w) run.x[w] = c.chunk(w)[tid];
    Vec<Tr<SEM>::VW> o[8];
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      run = OP::comb(x[i], run);
      o[i] = OP::value(run);
    }
    store_val8<SEM>(c, in.dst, tid, o);
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
  PYSTL_HD void step(Vec<1> l, Vec<1> r, bool cand, const RP&, float, float) {
    pm = fminf(pm, l.x[0]);
    if (cand) best = fmaxf(best, fminf(pm, r.x[0]));
  }
  PYSTL_HD Vec<1> done(int, const RP&) { return Vec<1>{{best}}; }
};

template <>
struct UntilState<SEM_CLASSIC_SRC> {  // classic.py:63-67 with provenance (earlier operand wins ties)
  Vec<2> pm, best;
  PYSTL_HD void init() { pm = Vec<2>{{f_inf(), 0.f}}; best = Vec<2>{{-f_inf(), 0.f}}; }
  PYSTL_HD void step(Vec<2> l, Vec<2> r, bool cand, const RP&, float, float) {
    pm = Red<SEM_CLASSIC_SRC, MODE_AND>::comb(pm, l);
    if (cand) best = Red<SEM_CLASSIC_SRC, MODE_OR>::comb(best, Red<SEM_CLASSIC_SRC, MODE_AND>::comb(pm, r));
  }
  PYSTL_HD Vec<2> done(int, const RP&) { return best; }
};

template <>
struct UntilState<SEM_CUMULATIVE> {  // cumulative.py:96-102
  float pm0, pm1, s0, s1;
  PYSTL_HD void init() { pm0 = pm1 = f_inf(); s0 = s1 = 0.f; }
  PYSTL_HD void step(Vec<2> l, Vec<2> r, bool cand, const RP&, float, float) {
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
  PYSTL_HD void step(Vec<1> l, Vec<1> r, bool cand, const RP& rp, float, float) {
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
  PYSTL_HD void step(Vec<1> l, Vec<1> r, bool cand, const RP& rp, float w0, float w1) {
    using RA = Red<SEM_AGM, MODE_AND>;
    using RO = Red<SEM_AGM, MODE_OR>;
    pre = RA::comb(pre, RA::lift(l, rp, 1.f));
    ++k;
    if (cand) {
      const Vec<1> y = RA::fin(pre, (float)k, rp);
      const Vec<2> pr = RA::comb(RA::lift(y, rp, w0), RA::lift(r, rp, w1));
      orr = RO::comb(orr, RO::lift(RA::fin(pr, w0 + w1, rp), rp, 1.f));
    }
  }
  PYSTL_HD Vec<1> done(int ncand, const RP& rp) { return Red<SEM_AGM, MODE_OR>::fin(orr, (float)ncand, rp); }
};

template <>
struct UntilState<SEM_DGMSR> {  // dgmsr.py:88-118 (uniform weights_left / weights_right)
  Vec<2> pre, orr;
  int k;
  PYSTL_HD void init() { pre = Red<SEM_DGMSR, MODE_AND>::ident(); orr = Red<SEM_DGMSR, MODE_OR>::ident(); k = 0; }
  PYSTL_HD void step(Vec<1> l, Vec<1> r, bool cand, const RP& rp, float w0, float w1) {
    using RA = Red<SEM_DGMSR, MODE_AND>;
    using RO = Red<SEM_DGMSR, MODE_OR>;
    pre = RA::comb(pre, RA::lift(l, rp, 1.f));
    ++k;
    if (cand) {
      const Vec<1> y = RA::fin(pre, (float)k, rp);
      const Vec<2> pr = RA::comb(RA::lift(y, rp, w0), RA::lift(r, rp, w1));
      orr = RO::comb(orr, RO::lift(RA::fin(pr, w0 + w1, rp), rp, 1.f));
    }
  }
  PYSTL_HD Vec<1> done(int ncand, const RP& rp) { return Red<SEM_DGMSR, MODE_OR>::fin(orr, (float)ncand, rp); }
};

template <int SEM, class EX>
PYSTL_HD void op_until(EX& ex, const Ctx& c, const TInstr& ins) {
  constexpr int VW = Tr<SEM>::VW;
  using V = Vec<VW>;
  const int nlive = c.nlive;
  const Pair in = decode_pair(ins);
  const int ua = ins.a, ub = ins.b;
  const bool open = ins.unbounded != 0;
  const float w0 = ins.w[0], w1 = ins.w[1];
  RP rp = c.rp;
  if constexpr (SEM == SEM_SMOOTH) {
    // shift = max over both operands of the exponent -v/tau
    rp.shift = fmaxf(smooth_shift<MODE_AND>(c, in.sl, in.nl), smooth_shift<MODE_AND>(c, in.sr, in.nr));
    // softmin over <= b+2 terms inside a softmax over <= b+1 candidates
    const float llo = in.nl ? -c.bounds()[2 * in.sl + 1] : c.bounds()[2 * in.sl];
    const float lhi = in.nl ? -c.bounds()[2 * in.sl] : c.bounds()[2 * in.sl + 1];
    const float rlo = in.nr ? -c.bounds()[2 * in.sr + 1] : c.bounds()[2 * in.sr];
    const float rhi = in.nr ? -c.bounds()[2 * in.sr] : c.bounds()[2 * in.sr + 1];
    const float nn = (float)((open ? c.lext : ub) + 2);
    const float pad = 2.f * c.rp.invk * fast_lg2(nn) + 1e-6f * (fabsf(llo) + fabsf(rhi));
    c.bounds()[2 * in.dst] = fminf(llo, rlo) - pad;
    c.bounds()[2 * in.dst + 1] = fmaxf(lhi, rhi) + pad;
  }
  // stage l and r (negations applied) in scratch planes so that phase 2 reads plain runs
  PYSTL_FOR_THREADS(ex, tid) {
    V l[8], r[8];
    load_val8<SEM>(c, in.sl, in.nl, tid, l);
    load_val8<SEM>(c, in.sr, in.nr, tid, r);
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
    const int b = open ? c.lext : ub;
    UntilState<SEM> st[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) st[i].init();
    int q_hi = p0 + 7 + b;
    if (q_hi > nlive - 1) q_hi = nlive - 1;
#pragma unroll 1
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
          if (kk >= 0 && kk <= b) st[i].step(l, r, kk >= ua, rp, w0, w1);
        }
      }
    }
    V o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int last = nlive - 1 - (p0 + i);  // min(b, T-1-t): api.py:284-293 + classic.py:59
      if (!open && last > ub) last = ub;
      const int ncand = last - ua + 1;
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
              if (kk >= ua) {
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
    case T_WIN_VH:
      if (in.mode == MODE_AND) op_win<SEM, MODE_AND>(ex, c, in, in.op == T_WIN_VH);
      else if (in.mode == MODE_OR) op_win<SEM, MODE_OR>(ex, c, in, in.op == T_WIN_VH);
      else if constexpr (has_sum) op_win<SEM, MODE_SUM>(ex, c, in, in.op == T_WIN_VH);
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
      if constexpr (SEM == SEM_CLASSIC || SEM == SEM_CLASSIC_SRC) op_clamp_scan<SEM>(ex, c, in);
      break;
    default:
      break;
  }
}

// Evaluate the whole program on a tile whose raw samples are already in shared memory.
template <int SEM, class EX>
PYSTL_HD void run_program(EX& ex, const Ctx& c, const TileProgram& P, const float* raw) {
  eval_predicates<SEM>(ex, c, P, raw);
  ex.sync();
#pragma unroll 1
  for (int i = 0; i < P.n_instr; ++i) run_instr<SEM>(ex, c, P.ins[i]);
}

// ------------------------------------------------------------------ result
// Write the formula value of this tile's outputs [0, n_out) to out_tile[p * C + w]
// (out_tile = &out[b, e0, 0]).  Outputs the reference raises on (api.py:218-221) are NaN.
// The root plane is re-read with lane-consecutive float4 ownership (float4 index
// tid + nt*j) so that the global stores are fully coalesced 128-bit stores.
// Optionally accumulates the sum of the valid outputs (all channels) for the fused
// mean-robustness reduction.
template <int SEM, class EX>
PYSTL_HD void store_root(EX& ex, const Ctx& c, const TileProgram& P, float* out_tile, int n_out, double* mean_acc) {
  constexpr int VW = Tr<SEM>::VW;
  const int root = P.root;
  const bool neg = P.root_neg != 0;
  int n_store = n_out < c.nlive ? n_out : c.nlive;
  const bool vec = (reinterpret_cast<uintptr_t>(out_tile) & 15) == 0;
  ex.sync();  // the root plane was written with blocked ownership
  PYSTL_FOR_THREADS(ex, tid) {
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int f = tid + c.nt * j;  // logical float4 index
      const int p = 4 * f;
      if (p >= n_store) continue;
      Vec<VW> v[4];
      {
        F4 q[VW];
#pragma unroll
        for (int w = 0; w < VW; ++w) q[w] = reinterpret_cast<const F4*>(c.plane(root, w, VW))[pf4(f)];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int w = 0; w < VW; ++w) v[k].x[w] = q[w].v[k];
          if (neg) v[k] = negate<SEM>(v[k]);
          if (p + k > c.last_ok) {
#pragma unroll
            for (int w = 0; w < VW; ++w) v[k].x[w] = f_nan();
          } else if (p + k < n_store) {
#pragma unroll
            for (int w = 0; w < VW; ++w) part += v[k].x[w];
          }
        }
      }
      float* dst = out_tile + (int64_t)p * VW;
      if (vec && p + 4 <= n_store) {
        F4* d4 = reinterpret_cast<F4*>(dst);
        if constexpr (VW == 1) {
          d4[0] = F4{{v[0].x[0], v[1].x[0], v[2].x[0], v[3].x[0]}};
        } else {
          d4[0] = F4{{v[0].x[0], v[0].x[1], v[1].x[0], v[1].x[1]}};
          d4[1] = F4{{v[2].x[0], v[2].x[1], v[3].x[0], v[3].x[1]}};
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (p + k < n_store) {
#pragma unroll
            for (int w = 0; w < VW; ++w) dst[k * VW + w] = v[k].x[w];
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
  float get(int tid) const { return stash[tid]; }
  float stash2[kMaxThreads];
  void put2(int tid, float a, float b) {
    stash[tid] = a;
    stash2[tid] = b;
  }
  void all_max2(float*, float& a, float& b) {
    a = b = -INFINITY;
    for (int i = 0; i < nt_; ++i) {
      a = fmaxf(a, stash[i]);
      b = fmaxf(b, stash2[i]);
    }
  }
  void atomic_add(float* p, float v) { *p += v; }
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
  // a[w*stride + c] <- comb(a[c+1], ..., a[n-1]) (identity for the last chunk), in time order
  template <int AW, class OP>
  void excl_suffix_scan(float* a, int n, int stride, float*) {
    typename OP::A run = OP::ident();
    for (int ch = n - 1; ch >= 0; --ch) {
      typename OP::A cur;
      for (int w = 0; w < AW; ++w) cur.x[w] = a[w * stride + ch];
      for (int w = 0; w < AW; ++w) a[w * stride + ch] = run.x[w];
      run = OP::comb(cur, run);
    }
  }
};
#endif

}  // namespace pystl
