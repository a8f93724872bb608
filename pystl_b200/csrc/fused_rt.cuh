// Runtime of the FUSED tier: register-resident tile kernels, specialised per formula.
//
// The interpreter tier (tile_eval.h) runs a lowered formula instruction by instruction and keeps
// every node value in a shared-memory plane.  The fused tier runs the same lowered program as
// straight-line code: codegen.cpp emits a short kernel body per (formula, CTA size, D) that calls
// the templates below with every window offset, operand sign and plane offset as a compile-time
// constant, so that
//   * a thread keeps the 8 consecutive samples it owns of EVERY node value in registers for the
//     whole kernel -- predicates, And / Or / Not never touch shared memory;
//   * only what other threads must see goes through shared memory: the chunk suffix / prefix
//     scans and chunk aggregates of the van Herk / Gil-Werman windows (api.py:215-253);
//   * windows of the same nesting depth share one barrier;
//   * chunk geometry (which aggregates lie inside the window of output i, where the partial
//     chunks start) folds to constants: no selects, no index arithmetic in the inner loops.
// The glue is compiled at run time by NVRTC (jit.cpp) or ahead of time into the cubin cache.
//
// Reference semantics restated: see the header of tile_eval.h (the reducers Red<SEM, MODE> are
// shared with the interpreter tier).
//
// This header is written in plain SIMT style.  The CPU test tier compiles it with g++ against
// tests/emu/simt_host.h (fibers); nothing in the product loads that build.
#pragma once
#include "tile_bwd.h"

#if defined(__CUDACC__)
#define PF_DEV __device__ __forceinline__
#define PF_COLD __device__ __noinline__  // rare exact paths: kept out of the hot instruction stream (scalars in, scalar out)
#define PF_KERNEL(NT, MINB) extern "C" __global__ void __launch_bounds__(NT, MINB)
#define PF_GRID_CONSTANT __grid_constant__
#define PF_DYN_SMEM(name) extern __shared__ __align__(1024) float name[]
namespace simt {
PF_DEV int tid() { return (int)threadIdx.x; }
PF_DEV int bid_x() { return (int)blockIdx.x; }
PF_DEV int bid_y() { return (int)blockIdx.y; }
PF_DEV void sync() { __syncthreads(); }
PF_DEV void syncwarp() { __syncwarp(); }
PF_DEV float shfl_down(float v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
PF_DEV float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
PF_DEV int redux_max_i(int v) { return __reduce_max_sync(0xffffffffu, v); }
PF_DEV void atomic_add(float* p, float v) { atomicAdd(p, v); }
PF_DEV void atomic_add(double* p, double v) { atomicAdd(p, v); }
// shared -> global bulk copy by the TMA engine (SASS UBLKCP.G.S): `bytes` % 16 == 0, both addresses
// 16-byte aligned.  The caller has made the shared image visible to the async proxy
// (fence_async_smem by every writer, then a barrier); the issuing thread waits until the engine has
// read the image, so that the CTA may exit.
PF_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// mbarrier + 2-D TMA tensor load (SASS UTMALDG): one thread arms the barrier with the byte count
// and issues the copy, everybody waits on the phase bit
PF_DEV void mbar_init(void* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PF_DEV void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
PF_DEV void mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}
PF_DEV void tma_store_3d(const void* tmap, const void* src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"((unsigned)__cvta_generic_to_shared(src_smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
PF_DEV void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
PF_DEV void tma_load_2d(void* dst_smem, const void* tmap, int c0, int c1, void* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst_smem)),
               "l"(tmap), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
PF_DEV void bulk_store_and_wait(void* dst_global, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global),
               "r"((unsigned)__cvta_generic_to_shared(src_smem)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
}  // namespace simt
#else
#include "simt_host.h"
#define PF_DEV inline
#define PF_COLD inline
#define PF_KERNEL(NT, MINB) extern "C" void
#define PF_GRID_CONSTANT
#define PF_DYN_SMEM(name) float* name = simt::dyn_smem()
#endif

namespace pystl {
namespace fused {

typedef unsigned long long u64;

// ------------------------------------------------------------------ tile context (registers)
struct TC {
  float* smem;
  int tid;
  int p0;       // first tile position this thread owns (8 * tid)
  int nlive;    // tile positions that exist in the trace: min(T - origin, lext)
  int last_ok;  // last tile position whose output does not raise (api.py:218-221); < 0: none
  int p_lo;     // tile positions [p_lo, p_hi) are the ones this CTA writes to HBM
  int p_hi;
  bool interior;  // the trace goes on past this tile: windows that leave the tile belong to halo outputs nobody consumes
  long long b;       // trace
  long long origin;  // trace time of tile position 0
  RP rp;
};

// grid = (traces, tiles per trace): no division to find the tile.  Tile arithmetic is 32-bit: the host only plans
// the fused tier for T < 2^31 - 2^16 (fused_geometry); 64 bits are for the element offsets c.b * T.
template <int NT>
PF_DEV TC tile_begin(const TileProgram& P, const TileGeom& g, float* smem) {
  TC c;
  c.smem = smem;
  c.tid = simt::tid();
  c.p0 = 8 * c.tid;
  c.b = (long long)simt::bid_x();
  const int T = (int)g.T;
  const int e0 = simt::bid_y() * g.lout;
  const int origin = e0 > g.hh ? e0 - g.hh : 0;  // backward tiles carry a left halo
  c.origin = origin;
  const int live = T - origin;
  c.nlive = live < 8 * NT ? live : 8 * NT;
  c.interior = live > 8 * NT;
  // last tile position whose output exists: T - 1 - start_slack - origin, clamped to [-1, 2^30]
  int ok = T - 1 - origin;  // >= -1 ... (a tile always starts inside the trace; T == 0 launches nothing)
  ok = (P.raises_always || P.start_slack > ok) ? -1 : ok - P.start_slack;
  c.last_ok = ok > (1 << 30) ? (1 << 30) : ok;
  c.p_lo = e0 - origin;
  c.p_hi = c.p_lo + g.lout < c.nlive ? c.p_lo + g.lout : c.nlive;
  c.rp = P.rp;
  return c;
}

PF_DEV bool own_all_live(const TC& c) { return c.p0 + 8 <= c.nlive; }
// threads whose 8 samples all lie past the horizon only keep the shared planes consistent
PF_DEV bool alive(const TC& c) { return c.p0 < c.nlive; }

// ------------------------------------------------------------------ signal rows -> registers
// The thread's 8 rows are 8*D consecutive floats of the trace: 128-bit loads when they all
// exist and the address is 16-byte aligned, guarded scalar loads otherwise (rows past the
// horizon read as 0 and are never consumed: the lifts replace them by the identity).
// L2 prefetch of the rows the SAME tile of trace b + ahead will load: CTAs are scheduled in trace
// order, so a CTA that starts `ahead` traces later finds its rows in L2 instead of HBM
// (the load latency at the head of every CTA was the top stall reason in the ncu profiles).
PF_DEV void prefetch_l2(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
template <int D>
PF_DEV void prefetch_rows(const TC& c, const float* __restrict__ signal, long long B, long long T, int ahead) {
  if (ahead <= 0 || c.b + ahead >= B || !own_all_live(c)) return;
  const char* src = reinterpret_cast<const char*>(signal + (((c.b + ahead) * T + c.origin + c.p0) * (long long)D));
  constexpr int BYTES = 8 * D * 4;  // this thread's rows
  if constexpr (BYTES >= 128) {
#pragma unroll
    for (int k = 0; k < BYTES / 128; ++k) prefetch_l2(src + 128 * k);
  } else {
    // several threads share a 128-byte line: one of them prefetches it
    if ((c.tid % (128 / BYTES)) == 0) prefetch_l2(src);
  }
}

template <int D>
PF_DEV void load_rows(const TC& c, const float* __restrict__ signal, long long T, float (&r)[8][D]) {
  const float* src = signal + ((c.b * T + c.origin + c.p0) * (long long)D);
  if (own_all_live(c) && ((u64)src & 15ull) == 0ull) {
    const F4* s4 = reinterpret_cast<const F4*>(src);
#pragma unroll
    for (int k = 0; k < 2 * D; ++k) {
#if defined(__CUDA_ARCH__)
      const float4 q = __ldg(reinterpret_cast<const float4*>(s4) + k);
      const float t[4] = {q.x, q.y, q.z, q.w};
#else
      const F4 q = s4[k];
      const float* t = q.v;
#endif
#pragma unroll
      for (int e = 0; e < 4; ++e) r[(4 * k + e) / D][(4 * k + e) % D] = t[e];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool ok = c.p0 + i < c.nlive;
#pragma unroll
      for (int d = 0; d < D; ++d) r[i][d] = ok ? src[i * D + d] : 0.f;
    }
  }
}
// D == 2: a thread's 8 rows are 64 bytes, so a 128-bit load of a warp touches 16 lines where a
// coalesced one touches 4, and the LSU data pipe -- not HBM -- is what the Classic kernels saturate
// (ncu: 97 % on cfg1x).  When the launch qualifies (TileGeom::use_tma) the CTA's tile is fetched by ONE
// TMA tensor copy instead: the signal is described as rows of 64 bytes with SWIZZLE_64B, so that
// the per-thread 128-bit reads of the staged tile (thread t owns row t, chunk k of it lives at
// k ^ ((t >> 1) & 3)) are bank-conflict free: 4 wavefronts per instruction instead of 16, and no
// global-load instructions at all.  Rows past the end of the tensor arrive as zeros; rows past the end
// of the trace belong to the next trace and are never consumed (the lifts replace them).
// All threads call it (it contains a barrier).  stage: 1024-byte aligned, 8 * NT * D floats.
template <int D, int NT>
PF_DEV void load_rows_tma(const TC& c, const TileGeom& g, float* stage, float* bar, float (&r)[8][D]) {
  static_assert(D == 2 || D == 4 || D == 8, "rows of 64 bytes (D = 2) or 128 bytes (D = 4, 8)");
  // D = 4 / 8: rows of 128 bytes with SWIZZLE_128B (chunk ^= row & 7); a thread owns one row (D = 4,
  // conflict-free) or two (D = 8, two-way conflicts -- still half the wavefronts of the
  // load + transpose through shared memory it replaces, and no global-load instructions)
  constexpr int ROWB = D == 2 ? 64 : 128, RPT = 32 * D / ROWB, SWZ = D == 2 ? 3 : 7;
  constexpr int TROWS = NT * RPT, NBOX = (TROWS + 255) / 256, ROWS = TROWS / NBOX;  // a TMA box has at most 256 rows
  static_assert(TROWS % NBOX == 0, "the tile splits into equal boxes");
  if (c.tid == 0) simt::mbar_init(bar);
  simt::sync();
  if (c.tid == 0) {
    simt::mbar_expect_tx(bar, 8u * NT * D * 4u);
    const long long row = ((c.b * g.T + c.origin) * (long long)D * 4) / ROWB;
#pragma unroll
    for (int k = 0; k < NBOX; ++k) simt::tma_load_2d(stage + k * ROWS * (ROWB / 4), &g.tmap, 0, (int)row + k * ROWS, bar);
  }
  simt::mbar_wait(bar, 0u);
  const char* base = reinterpret_cast<const char*>(stage);
#pragma unroll
  for (int k = 0; k < 2 * D; ++k) {
    const unsigned o = (unsigned)c.tid * (32u * D) + 16u * k;  // linear byte offset of the 16-byte chunk
    const F4 q = *reinterpret_cast<const F4*>(base + (o ^ (((o >> 7) & SWZ) << 4)));
#pragma unroll
    for (int e = 0; e < 4; ++e) r[(4 * k + e) / D][(4 * k + e) % D] = q.v[e];
  }
}
// Wide rows (D >= 4): a thread's 8 rows are 128+ bytes apart from its neighbour's, so per-thread
// 128-bit loads touch 32 different lines per instruction and saturate the L1 pipeline (ncu: "Mem
// Busy" 90 % on cfg5).  Instead the warp loads its 32 * 8 * D contiguous floats with fully
// coalesced 128-bit loads and transposes them through its own slice of shared memory; the slice
// is swizzled so that both the coalesced stores and the per-thread row reads are conflict-free.
// Only the warp synchronises.  stage: this CTA's staging area (8 * NT * D floats).
template <int D>
PF_DEV void load_rows_coop(const TC& c, const float* __restrict__ signal, long long T, float* stage, float (&r)[8][D]) {
  constexpr int R = 2 * D;  // 16-byte granules per thread
  const int lane = c.tid & 31, w0 = c.tid & ~31;
  const float* src = signal + ((c.b * T + c.origin + 8 * w0) * (long long)D);
  const bool coop = 8 * w0 + 256 <= c.nlive && ((u64)src & 15ull) == 0ull;  // warp-uniform
  if (!coop) {
    load_rows<D>(c, signal, T, r);
    return;
  }
  F4* st4 = reinterpret_cast<F4*>(stage) + w0 * R;
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int gidx = k * 32 + lane;
#if defined(__CUDA_ARCH__)
    const float4 q = __ldg(reinterpret_cast<const float4*>(src) + gidx);
    const F4 f = F4{{q.x, q.y, q.z, q.w}};
#else
    const F4 f = reinterpret_cast<const F4*>(src)[gidx];
#endif
    st4[gidx ^ ((gidx >> (R <= 8 ? 3 : 4)) & 7)] = f;
  }
  simt::syncwarp();
#pragma unroll
  for (int m = 0; m < R; ++m) {
    const int gidx = R * lane + m;
    const F4 f = st4[gidx ^ ((gidx >> (R <= 8 ? 3 : 4)) & 7)];
#pragma unroll
    for (int e = 0; e < 4; ++e) r[(4 * m + e) / D][(4 * m + e) % D] = f.v[e];
  }
}

// one channel of a wide signal (D > 8): strided scalar loads
PF_DEV void load_channel(const TC& c, const float* __restrict__ signal, long long T, int D, int ch, float (&x)[8]) {
  const float* src = signal + ((c.b * T + c.origin + c.p0) * (long long)D) + ch;
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (c.p0 + i < c.nlive) ? src[(long long)i * D] : 0.f;
}

// ------------------------------------------------------------------ block reductions
PF_DEV float warp_max(float v) {  // one REDUX: floats mapped to integers of the same order
  int k = f2i(v);
  k = k >= 0 ? k : k ^ 0x7fffffff;
  k = simt::redux_max_i(k);
  k = k >= 0 ? k : k ^ 0x7fffffff;
  return i2f(k);
}

// in: per-thread maxima m[0..N) ; out: block-wide maxima in every thread.  red: N * NT/32 floats.
template <int NT, int N>
PF_DEV void block_max(const TC& c, float* red, float (&m)[N]) {
  constexpr int NW = NT / 32;
  const int lane = c.tid & 31, warp = c.tid >> 5;
#pragma unroll
  for (int q = 0; q < N; ++q) {
    const float v = warp_max(m[q]);
    if (lane == 0) red[q * NW + warp] = v;
  }
  simt::sync();
#pragma unroll
  for (int q = 0; q < N; ++q) {
    float a = red[q * NW];
#pragma unroll
    for (int w = 1; w < NW; ++w) a = fmaxf(a, red[q * NW + w]);
    m[q] = a;
  }
}

// max of +v (NEG: of -v) over the thread's live samples
template <bool NEG>
PF_DEV float max8(const TC& c, const Vec<1> (&v)[8]) {
  float m = -f_inf();
  if (own_all_live(c)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) m = fmaxf(m, NEG ? -v[i].x[0] : v[i].x[0]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (c.p0 + i < c.nlive) m = fmaxf(m, NEG ? -v[i].x[0] : v[i].x[0]);
  }
  return m;
}

// smooth: range of softmin / softmax over n terms of operands within [lo, hi]:
//   min - tau ln n <= softmin <= min,   max <= softmax <= max + tau ln n   (+ rounding slack)
template <int MODE>
PF_DEV float smooth_lo(const RP& rp, float lo, float n) {
  return MODE == MODE_AND ? lo - rp.invk * fast_lg2(n > 1.f ? n : 1.f) - 2e-6f * fabsf(lo) - 1e-30f : lo;
}
template <int MODE>
PF_DEV float smooth_hi(const RP& rp, float hi, float n) {
  return MODE == MODE_AND ? hi : hi + rp.invk * fast_lg2(n > 1.f ? n : 1.f) + 2e-6f * fabsf(hi) + 1e-30f;
}

// ------------------------------------------------------------------ static chunk geometry
template <int X>
struct FloorDiv {
  static constexpr int q8 = X >= 0 ? X / 8 : -((-X + 7) / 8);
  static constexpr int r8 = X - 8 * q8;
  static constexpr int q4 = X >= 0 ? X / 4 : -((-X + 3) / 4);
  static constexpr int r4 = X - 4 * q4;
};

// 8 consecutive logical samples starting at 8 * chunk + OFF of a swizzled plane
template <int OFF>
PF_DEV void load_run_s(const float* plane, int chunk, float (&v)[8]) {
  const F4* P4 = reinterpret_cast<const F4*>(plane);
  constexpr int q4 = FloorDiv<OFF>::q4, r4 = FloorDiv<OFF>::r4;
  const int f0 = 2 * chunk + q4;
  if constexpr (r4 == 0) {
    const F4 a = P4[pf4(f0)], b = P4[pf4(f0 + 1)];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = a.v[k];
      v[4 + k] = b.v[k];
    }
  } else {
    const F4 a = P4[pf4(f0)], b = P4[pf4(f0 + 1)], d = P4[pf4(f0 + 2)];
    const float t[12] = {a.v[0], a.v[1], a.v[2], a.v[3], b.v[0], b.v[1], b.v[2], b.v[3], d.v[0], d.v[1], d.v[2], d.v[3]};
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = t[i + r4];
  }
}

template <int AW>
PF_DEV void split_words(const Vec<AW> (&a)[8], int w, float (&t)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = a[i].x[w];
}

// ------------------------------------------------------------------ sliding window
// out[p] = (+)_{k in [WA, WB]} x[p + k] over the tile for an associative (+) = R::comb with
// identity R::ident(); operands are combined in time order.  WA <= WB, any signs (the adjoint
// windows of the backward pass look back).  phase1 / phase2 are separated by ONE barrier, which
// the generated kernel shares between all windows of the same depth.
//
//   W <= 8  direct: x is staged; every output folds its W samples with a doubling table
//   W >= 9  van Herk / Gil-Werman over the 8-sample chunks the threads own:
//           out[p] = SUF[p+WA] (+) AGG[chunks strictly inside] (+) PRE[p+WB]
//           (own registers replace SUF when WA == 0 and PRE when WB == 0)
template <class R, int WA, int WB, int NT>
struct WinOp {
  using A = typename R::A;
  static constexpr int AW = (int)(sizeof(A) / sizeof(float));
  static constexpr int LEXT = 8 * NT;
  static constexpr int W = WB - WA + 1;
  static constexpr bool VH = W > 8;
  static constexpr int GL = WA < 0 ? ((-WA + 7) / 8) * 8 : 0;
  static constexpr int GR = (WB > 0 ? ((WB + 7) / 8) * 8 : 0) + 8;
  static constexpr int PLANE = GL + LEXT + GR;
  static constexpr int NCHK = PLANE / 8;
  static constexpr bool USE_SUF = VH ? (WA != 0) : true;  // direct: plane 0 holds x itself
  static constexpr bool USE_PRE = VH && (WB != 0);
  static constexpr int NPL = (USE_SUF ? 1 : 0) + (USE_PRE ? 1 : 0);
  static constexpr int AGGN = VH ? ((NCHK + 3) / 4) * 4 : 0;
  static constexpr int FLOATS = AW * (NPL * PLANE + AGGN);

  PF_DEV static float* plane0(float* scr, int w) { return scr + (w * NPL) * PLANE + GL; }
  PF_DEV static float* plane1(float* scr, int w) { return scr + (w * NPL + (USE_SUF ? 1 : 0)) * PLANE + GL; }
  PF_DEV static float* agg(float* scr, int w) { return scr + AW * NPL * PLANE + w * AGGN + GL / 8; }

  // identity into the guard bands (chunks [-GL/8, 0) and [NT, NT + GR/8))
  PF_DEV static void fill_guards(float* scr, int tid) {
    const A id = R::ident();
    constexpr int NG = (GL + GR) / 8;
    for (int g = tid; g < NG; g += NT) {
      const int ch = g < GR / 8 ? NT + g : -(g - GR / 8) - 1;
#pragma unroll
      for (int w = 0; w < AW; ++w) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = id.x[w];
        if constexpr (USE_SUF) store8(plane0(scr, w), ch, t);
        if constexpr (USE_PRE) store8(plane1(scr, w), ch, t);
        if constexpr (VH) agg(scr, w)[ch] = id.x[w];
      }
    }
  }

  // x: the thread's lifted samples.  suf / pre: its chunk scans, kept in registers for phase2.
  PF_DEV static void phase1(float* scr, int tid, const A (&x)[8], A (&suf)[8], A (&pre)[8]) {
    if constexpr (!VH) {
#pragma unroll
      for (int i = 0; i < 8; ++i) suf[i] = pre[i] = x[i];
#pragma unroll
      for (int w = 0; w < AW; ++w) {
        float t[8];
        split_words<AW>(x, w, t);
        store8(plane0(scr, w), tid, t);
      }
    } else {
      pre[0] = x[0];
#pragma unroll
      for (int i = 1; i < 8; ++i) pre[i] = R::comb(pre[i - 1], x[i]);
      suf[7] = x[7];
#pragma unroll
      for (int i = 6; i >= 0; --i) suf[i] = R::comb(x[i], suf[i + 1]);
#pragma unroll
      for (int w = 0; w < AW; ++w) {
        float t[8];
        if constexpr (USE_SUF) {
          split_words<AW>(suf, w, t);
          store8(plane0(scr, w), tid, t);
        }
        if constexpr (USE_PRE) {
          split_words<AW>(pre, w, t);
          store8(plane1(scr, w), tid, t);
        }
        agg(scr, w)[tid] = pre[7].x[w];
      }
    }
  }
  // a thread past the horizon: its chunk is the identity
  PF_DEV static void phase1_dead(float* scr, int tid) {
    const A id = R::ident();
#pragma unroll
    for (int w = 0; w < AW; ++w) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = id.x[w];
      if constexpr (USE_SUF) store8(plane0(scr, w), tid, t);
      if constexpr (USE_PRE) store8(plane1(scr, w), tid, t);
      if constexpr (VH) agg(scr, w)[tid] = id.x[w];
    }
  }

  PF_DEV static A load_agg(const float* scr, int ch) {
    A r;
#pragma unroll
    for (int w = 0; w < AW; ++w) r.x[w] = agg(const_cast<float*>(scr), w)[ch];
    return r;
  }

  // fold LEN consecutive entries of e[] starting at S with the doubling tables y2 / y4 / y8
  template <int S, int LEN, int N>
  PF_DEV static A fold(const A (&e)[N], const A (&y2)[N], const A (&y4)[N], const A (&y8)[N]) {
    if constexpr (LEN == 1) return e[S];
    else if constexpr (LEN == 2) return y2[S];
    else if constexpr (LEN == 4) return y4[S];
    else if constexpr (LEN == 8) return y8[S];
    else if constexpr (LEN > 4) return R::comb(y4[S], fold<S + 4, LEN - 4, N>(e, y2, y4, y8));
    else return R::comb(y2[S], e[S + 2]);  // LEN == 3
  }

  PF_DEV static void phase2(float* scr, int tid, const A (&x)[8], const A (&suf)[8], const A (&pre)[8], A (&acc)[8]) {
    if constexpr (!VH) {
      // e[j] = x[p0 + WA + j], j in [0, W + 7): own samples from registers, the rest from the plane
      constexpr int N = W + 7;
      A e[N], y2[N], y4[N], y8[N];
      constexpr int f_lo = FloorDiv<WA>::q4, f_hi = FloorDiv<WB + 7>::q4;
#pragma unroll
      for (int f = f_lo; f <= f_hi; ++f) {
        if (f == 0 || f == 1) continue;  // own chunk
#pragma unroll
        for (int w = 0; w < AW; ++w) {
          const F4 q = reinterpret_cast<const F4*>(plane0(scr, w))[pf4(2 * tid + f)];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = 4 * f + k - WA;
            if (j >= 0 && j < N) e[j].x[w] = q.v[k];
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = i - WA;
        if (j >= 0 && j < N) e[j] = x[i];
      }
      if constexpr (W >= 2) {
#pragma unroll
        for (int j = 0; j + 1 < N; ++j) y2[j] = R::comb(e[j], e[j + 1]);
      }
      if constexpr (W >= 4) {
#pragma unroll
        for (int j = 0; j + 3 < N; ++j) y4[j] = R::comb(y2[j], y2[j + 2]);
      }
      if constexpr (W >= 8) {
#pragma unroll
        for (int j = 0; j + 7 < N; ++j) y8[j] = R::comb(y4[j], y4[j + 4]);
      }
      fold_all(e, y2, y4, y8, acc);
      return;
    } else {
      constexpr int qa = FloorDiv<WA>::q8, ra = FloorDiv<WA>::r8, qb = FloorDiv<WB>::q8, rb = FloorDiv<WB>::r8;
      A sv[8], ev[8];
      if constexpr (WA == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sv[i] = suf[i];
      } else {
#pragma unroll
        for (int w = 0; w < AW; ++w) {
          float t[8];
          load_run_s<WA>(plane0(scr, w), tid, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) sv[i].x[w] = t[i];
        }
      }
      if constexpr (WB == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ev[i] = pre[i];
      } else {
#pragma unroll
        for (int w = 0; w < AW; ++w) {
          float t[8];
          load_run_s<WB>(plane1(scr, w), tid, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) ev[i].x[w] = t[i];
        }
      }
      // whole chunks strictly between the two partial ones: chunk tid + k for k in [lo_i, hi_i],
      // lo_i = qa + ((ra + i) >> 3) + 1, hi_i = qb + ((rb + i) >> 3) - 1.  All i share the core
      // [qa + 2, qb - 1]; chunk qa + 1 and chunk qb join it for some i only.
      constexpr int NCORE = qb - qa - 2;
      A core = R::ident();
      if constexpr (NCORE > 0) {
        core = load_agg(scr, tid + qa + 2);
#pragma unroll
        for (int k = 1; k < NCORE; ++k) core = R::comb(core, load_agg(scr, tid + qa + 2 + k));
      }
      const A agl = load_agg(scr, tid + qa + 1);
      const A agr = load_agg(scr, tid + qb);
      // the (at most four) distinct sets of whole chunks, in time order
      const A m_l = NCORE > 0 ? R::comb(agl, core) : agl;
      const A m_r = NCORE > 0 ? R::comb(core, agr) : agr;
      const A m_lr = R::comb(m_l, agr);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int lo = qa + ((ra + i) >> 3) + 1, hi = qb + ((rb + i) >> 3) - 1;
        const bool inc_l = (lo == qa + 1) && (qa + 1 <= hi);
        const bool inc_r = (hi == qb) && (lo <= qb) && (qb != qa + 1 || !inc_l);
        if (inc_l && inc_r) acc[i] = R::comb(R::comb(sv[i], m_lr), ev[i]);
        else if (inc_l) acc[i] = R::comb(R::comb(sv[i], m_l), ev[i]);
        else if (inc_r) acc[i] = R::comb(R::comb(sv[i], m_r), ev[i]);
        else if (NCORE > 0) acc[i] = R::comb(R::comb(sv[i], core), ev[i]);
        else acc[i] = R::comb(sv[i], ev[i]);
      }
    }
  }

  template <int N>
  PF_DEV static void fold_all(const A (&e)[N], const A (&y2)[N], const A (&y4)[N], const A (&y8)[N], A (&acc)[8]) {
    acc[0] = fold<0, W, N>(e, y2, y4, y8);
    acc[1] = fold<1, W, N>(e, y2, y4, y8);
    acc[2] = fold<2, W, N>(e, y2, y4, y8);
    acc[3] = fold<3, W, N>(e, y2, y4, y8);
    acc[4] = fold<4, W, N>(e, y2, y4, y8);
    acc[5] = fold<5, W, N>(e, y2, y4, y8);
    acc[6] = fold<6, W, N>(e, y2, y4, y8);
    acc[7] = fold<7, W, N>(e, y2, y4, y8);
  }
};

template <int VW>
PF_DEV void stash8(float* plane, int tid, const Vec<VW> (&v)[8], int w) {
  float t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = v[i].x[w];
  store8(plane, tid, t);
}

// ------------------------------------------------------------------ classic unbounded until
// W(s) = min(l_s, max(r_s, W(s+1))): a suffix scan of the clamp functions x -> min(A, max(B, x))
// over the tile (tile_eval.h: ClampOp).  Chunk-local composition in
// registers, warp suffix scan with shuffles, warp totals through shared memory: one barrier.
// Every thread of the CTA calls both phases (threads past the horizon pass the identity).
template <int SEM, int NT>
struct ClampScanOp {
  using OP = ClampOp<SEM>;
  using A = typename OP::A;
  static constexpr int W = OP::W;
  static constexpr int NW = NT / 32;
  static constexpr int FLOATS = W * 32;

  PF_DEV static void phase1(float* scr, const TC& c, const A (&x)[8], A& excl) {
    A v = x[7];
#pragma unroll
    for (int i = 6; i >= 0; --i) v = OP::comb(x[i], v);
    const int lane = c.tid & 31, warp = c.tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      A other;
#pragma unroll
      for (int w = 0; w < W; ++w) other.x[w] = simt::shfl_down(v.x[w], o);
      if (lane + o < 32) v = OP::comb(v, other);
    }
    if (lane == 0) {
#pragma unroll
      for (int w = 0; w < W; ++w) scr[w * 32 + warp] = v.x[w];
    }
#pragma unroll
    for (int w = 0; w < W; ++w) excl.x[w] = simt::shfl_down(v.x[w], 1);
    if (lane == 31) excl = OP::ident();
  }

  template <int VW>
  PF_DEV static void phase2(const float* scr, const TC& c, const A (&x)[8], A excl, Vec<VW> (&o)[8]) {
    const int warp = c.tid >> 5;
    A tail = OP::ident();
    for (int ww = NW - 1; ww > warp; --ww) {
      A t;
#pragma unroll
      for (int w = 0; w < W; ++w) t.x[w] = scr[w * 32 + ww];
      tail = OP::comb(t, tail);
    }
    A run = OP::comb(excl, tail);
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      run = OP::comb(x[i], run);
      o[i] = OP::value(run);
    }
  }
};

// smooth until: one output recomputed exactly in the log2 domain (rare path of UntilOp::smooth_fix)
template <int UA>
PF_COLD float smooth_until_fix_one(int p, int last, float k, float invk, const float* pl, const float* pr) {
  float mS = -f_inf(), S = 0.f, mA = -f_inf(), A = 0.f;
  for (int kk = 0; kk <= last; ++kk) {
    const float zl = -pl[pidx(p + kk)] * k;
    if (zl > mS) {
      S = S * fast_ex2(mS - zl) + 1.f;
      mS = zl;
    } else {
      S += fast_ex2(zl - mS);
    }
    if (kk >= UA) {
      const float zr = -pr[pidx(p + kk)] * k;
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
  return (mA + fast_lg2(A)) * invk;
}

template <bool B>
struct BoolC {
  static constexpr bool value = B;
};

// Two reciprocals from ONE MUFU: 1 / a = b / (a b), 1 / b = a / (a b).  The smooth Until walks are bound by the XU pipe (one
// rcp per candidate, ncu: 83 %); this trades every second rcp for three multiplies on the FMA pipe.  The product must stay
// normal: the walks call an output healthy only while its sum of terms is below kSmoothUntilMax = 2^60, i.e. while each of
// its denominators is above 2^-60; a product that underflows makes both terms inf, which sends both outputs to the exact path.
constexpr float kSmoothUntilMax = 1.15292150e18f;  // 2^60
PF_DEV void rcp2(float a, float b, float& ra, float& rb) {
  const float r = fast_rcp(a * b);
  ra = r * b;
  rb = r * a;
}

// ------------------------------------------------------------------ until, O(b) per sample
// OR_{i in [UA, UB]} AND(AND_{k<=i} l[t+k], r[t+i]) (api.py:284-302 + each semantics'
// temporal_until) with the per-semantics running state of tile_eval.h (UntilState): 8 outputs
// per thread are fed from 128-bit runs of the staged operands, no O(b^2) prefix recomputation.
// phase1 stages both operands (negations applied); one barrier; phase2 walks the window.
// EXPL (Smooth, forward kernels): two more planes hold lambda = 2^(-k l - shift), rho = 2^(-k r - shift), exponentiated ONCE per
// sample by its owner; the walk -- which visits every position from about (UB + 8) / 8 threads -- then costs no exponential
// (cfg5 Smooth forward: 32 -> 22 MUFU operations per sample on a kernel whose XU pipe was 74-83 % busy).
#ifndef PYSTL_DGMSR_SELECT
#define PYSTL_DGMSR_SELECT 1
#endif
template <int SEM, int UA, int UB, int NT, bool EXPL = false>
struct UntilOp {
  static constexpr int VW = Tr<SEM>::VW;
  using V = Vec<VW>;
  static constexpr int LEXT = 8 * NT;
  static constexpr int FLOATS = (2 * VW + (EXPL ? 2 : 0)) * LEXT;
  static_assert(!EXPL || SEM == SEM_SMOOTH, "exponential planes belong to the smooth walk");
  PF_DEV static float* lplane(float* scr, int w) { return scr + w * LEXT; }
  PF_DEV static float* rplane(float* scr, int w) { return scr + (VW + w) * LEXT; }
  PF_DEV static float* elplane(float* scr) { return scr + 2 * VW * LEXT; }
  PF_DEV static float* erplane(float* scr) { return scr + (2 * VW + 1) * LEXT; }

  PF_DEV static void phase1(float* scr, int tid, const V (&l)[8], const V (&r)[8]) {
#pragma unroll
    for (int w = 0; w < VW; ++w) {
      stash8<VW>(lplane(scr, w), tid, l, w);
      stash8<VW>(rplane(scr, w), tid, r, w);
    }
  }
  PF_DEV static void phase1(float* scr, int tid, const V (&l)[8], const V (&r)[8], const RP& rp) {
    phase1(scr, tid, l, r);
    if constexpr (EXPL) {
      const float nk = -rp.k, ns = -rp.shift;
      Vec<1> el[8], er[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        el[i].x[0] = fast_ex2(fmaf(l[i].x[0], nk, ns));
        er[i].x[0] = fast_ex2(fmaf(r[i].x[0], nk, ns));
      }
      stash8<1>(elplane(scr), tid, el, 0);
      stash8<1>(erplane(scr), tid, er, 0);
    }
  }

  // number of candidates of output p: min(UB, T-1-t) - UA + 1 (api.py:284-293 + classic.py:59)
  PF_DEV static int ncand(const TC& c, int p) {
    int last = c.nlive - 1 - p;
    if (last > UB) last = UB;
    return last - UA + 1;
  }

  // Cumulative fast path (cumulative.py:96-103: U = sum_i min(r_i, min_{k<=i} l_k) per channel).
  // Past the thread's own chunk the 8 outputs share everything: for q >= p0 + 8
  //     min_{k in [p0+i, q]} l_k = min(h_i, M_q),   h_i = min l[p0+i .. p0+7],   M_q = min l[p0+8 .. q],
  // so the candidate of output i at q is min(G_q, h_i) with G_q = min(r_q, M_q): two mins per
  // position for the thread, one min + one add per (position, output), operands read once per
  // thread (128-bit runs) instead of a guarded state update per (position, output).
  // On an interior tile (the trace goes on past it) the outputs whose windows leave the tile are never
  // consumed and the loads are simply clamped.  On the last tile of a trace the windows are clipped at
  // the horizon (api.py:49-54): positions past kmax = nlive - 1 - p0 are skipped, outputs without a
  // candidate are NaN (the reference raises).
  PF_DEV static int clip_kmax(const TC& c) { return c.interior ? 0x3fffffff : c.nlive - 1 - c.p0; }
  PF_DEV static bool cum_fast_ok(const TC&) { return true; }
  // (a second instantiation without the per-position horizon tests for the threads no window of which is clipped was
  // measured: cfg4 4.58 -> 4.51 ms, but cfg5 2.74 -> 3.19 ms -- two copies of the walk cost more than the tests)
  PF_DEV static void phase2_cum(float* scr, const TC& c, V (&o)[8]) {
    static_assert(UA >= 0 && UB >= UA, "bounded until");
    const int f0 = c.p0 >> 2;
    const int kmax = clip_kmax(c);
    constexpr int FMAX = LEXT / 4 - 1;
    constexpr int KS0 = UA + 7 > 8 ? UA + 7 : 8;       // first offset at which every output has a candidate
    constexpr int KS = (KS0 + 3) & ~3;                 // ... rounded up to a 128-bit group
    constexpr int KE = ((UB + 1) & ~3) > KS ? ((UB + 1) & ~3) : KS;  // body = [KS, KE), every output active
#pragma unroll
    for (int w = 0; w < VW; ++w) {
      const F4* L4 = reinterpret_cast<const F4*>(lplane(scr, w));
      const F4* R4 = reinterpret_cast<const F4*>(rplane(scr, w));
      float acc[8], h[8];
      {  // own chunk: triangular
        const F4 la = L4[pf4(f0)], lb = L4[pf4(f0 + 1)], ra = R4[pf4(f0)], rb = R4[pf4(f0 + 1)];
        const float lo[8] = {la.v[0], la.v[1], la.v[2], la.v[3], lb.v[0], lb.v[1], lb.v[2], lb.v[3]};
        const float ro[8] = {ra.v[0], ra.v[1], ra.v[2], ra.v[3], rb.v[0], rb.v[1], rb.v[2], rb.v[3]};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float pm = f_inf(), a = 0.f;
#pragma unroll
          for (int q = i; q < 8; ++q) {
            pm = fminf(pm, lo[q]);
            if (q - i >= UA && q - i <= UB && q <= kmax) a += fminf(ro[q], pm);
          }
          acc[i] = a;
          h[i] = pm;
        }
      }
      float M = f_inf();
      // one position at offset kq (compile-time) from p0: value pair (l, r)
#define PF_CUM_STEP(KQ, LV, RV)                                            \
  if ((KQ) <= kmax) {                                                      \
    M = fminf(M, (LV));                                                    \
    const float G = fminf((RV), M);                                        \
    _Pragma("unroll") for (int i = 0; i < 8; ++i) {                        \
      if ((KQ) - i >= UA && (KQ) - i <= UB) acc[i] += fminf(G, h[i]);      \
    }                                                                      \
  }
      // head [8, KS): static conditions (offsets below UA only extend the running minimum)
#pragma unroll
      for (int g = 2; g < KS / 4; ++g) {
        const int fc = (f0 + g) < FMAX ? (f0 + g) : FMAX;
        const F4 ql = L4[pf4(fc)], qr = R4[pf4(fc)];
#pragma unroll
        for (int k = 0; k < 4; ++k) PF_CUM_STEP(4 * g + k, ql.v[k], qr.v[k])
      }
      // body [KS, KE): every output takes every position
#pragma unroll 2
      for (int g = KS / 4; g < KE / 4; ++g) {
        const int fc = (f0 + g) < FMAX ? (f0 + g) : FMAX;
        const F4 ql = L4[pf4(fc)], qr = R4[pf4(fc)];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (4 * g + k > kmax) break;
          M = fminf(M, ql.v[k]);
          const float G = fminf(qr.v[k], M);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += fminf(G, h[i]);
        }
      }
      // tail [KE, UB + 8): static conditions
#pragma unroll
      for (int g = KE / 4; g <= (UB + 7) / 4; ++g) {
        const int fc = (f0 + g) < FMAX ? (f0 + g) : FMAX;
        const F4 ql = L4[pf4(fc)], qr = R4[pf4(fc)];
#pragma unroll
        for (int k = 0; k < 4; ++k) PF_CUM_STEP(4 * g + k, ql.v[k], qr.v[k])
      }
#undef PF_CUM_STEP
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i].x[w] = (kmax - i < UA) ? f_nan() : acc[i];
    }
  }

  // Smooth fast path (smooth.py:95-102 through SURVEY.md appendix A4): U = tau ln sum_i 1 / (S_i + rho_i),
  // S_i = sum_{k<=i} lambda_k.  Same sharing as above with sums in place of minima: past the own
  // chunk S_i(q) = hS_i + M_q, so lambda_q and rho_q are exponentiated once per thread and position
  // (2 MUFU) and every candidate costs one reciprocal instead of two exponentials and a division.
  PF_DEV static void phase2_smooth(float* scr, const TC& c, const RP& rp, V (&o)[8], UntilState<SEM> (&st)[8]) {
    const int f0 = c.p0 >> 2;
    constexpr int FMAX = LEXT / 4 - 1;
    constexpr int KS0 = UA + 7 > 8 ? UA + 7 : 8;
    constexpr int KS = (KS0 + 3) & ~3;
    constexpr int KE = ((UB + 1) & ~3) > KS ? ((UB + 1) & ~3) : KS;
    const F4* L4 = reinterpret_cast<const F4*>(EXPL ? elplane(scr) : lplane(scr, 0));
    const F4* R4 = reinterpret_cast<const F4*>(EXPL ? erplane(scr) : rplane(scr, 0));
    const float nk = -rp.k, ns = -rp.shift;
    // lambda_q / rho_q of a staged pair: read as they stand from the exponential planes, or exponentiated here
    auto ex = [&](float v) { return EXPL ? v : fast_ex2(fmaf(v, nk, ns)); };
    const int kmax = clip_kmax(c);
    float acc[8], h[8];
    {
      const F4 la = L4[pf4(f0)], lb = L4[pf4(f0 + 1)], ra = R4[pf4(f0)], rb = R4[pf4(f0 + 1)];
      float lo[8] = {la.v[0], la.v[1], la.v[2], la.v[3], lb.v[0], lb.v[1], lb.v[2], lb.v[3]};
      float ro[8] = {ra.v[0], ra.v[1], ra.v[2], ra.v[3], rb.v[0], rb.v[1], rb.v[2], rb.v[3]};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        lo[q] = ex(lo[q]);
        ro[q] = ex(ro[q]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float S = 0.f, a = 0.f;
#pragma unroll
        for (int q = i; q < 8; ++q) {
          S += lo[q];
          if (q - i >= UA && q - i <= UB && q <= kmax) a += fast_rcp(S + ro[q]);
        }
        acc[i] = a;
        h[i] = S;
      }
    }
    float M = 0.f;
#define PF_SM_STEP(KQ, LV, RV)                                                   \
  if ((KQ) <= kmax) {                                                            \
    M += ex(LV);                                                                 \
    const float G = M + ex(RV);                                                  \
    _Pragma("unroll") for (int i = 0; i < 8; i += 2) {                           \
      const bool c0 = (KQ) - i >= UA && (KQ) - i <= UB;                          \
      const bool c1 = (KQ) - i - 1 >= UA && (KQ) - i - 1 <= UB;                  \
      if (c0 && c1) {                                                            \
        float t0, t1;                                                            \
        rcp2(G + h[i], G + h[i + 1], t0, t1);                                    \
        acc[i] += t0;                                                            \
        acc[i + 1] += t1;                                                        \
      } else if (c0) {                                                           \
        acc[i] += fast_rcp(G + h[i]);                                            \
      } else if (c1) {                                                           \
        acc[i + 1] += fast_rcp(G + h[i + 1]);                                    \
      }                                                                          \
    }                                                                            \
  }
#pragma unroll
    for (int g = 2; g < KS / 4; ++g) {
      const int fc = (f0 + g) < FMAX ? (f0 + g) : FMAX;
      const F4 ql = L4[pf4(fc)], qr = R4[pf4(fc)];
#pragma unroll
      for (int k = 0; k < 4; ++k) PF_SM_STEP(4 * g + k, ql.v[k], qr.v[k])
    }
#pragma unroll 2
    for (int g = KS / 4; g < KE / 4; ++g) {
      const int fc = (f0 + g) < FMAX ? (f0 + g) : FMAX;
      const F4 ql = L4[pf4(fc)], qr = R4[pf4(fc)];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (4 * g + k > kmax) break;
        M += ex(ql.v[k]);
        const float G = M + ex(qr.v[k]);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          float t0, t1;
          rcp2(G + h[i], G + h[i + 1], t0, t1);
          acc[i] += t0;
          acc[i + 1] += t1;
        }
      }
    }
#pragma unroll
    for (int g = KE / 4; g <= (UB + 7) / 4; ++g) {
      const int fc = (f0 + g) < FMAX ? (f0 + g) : FMAX;
      const F4 ql = L4[pf4(fc)], qr = R4[pf4(fc)];
#pragma unroll
      for (int k = 0; k < 4; ++k) PF_SM_STEP(4 * g + k, ql.v[k], qr.v[k])
    }
#undef PF_SM_STEP
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      st[i].S = 0.f;
      st[i].acc = acc[i];  // (smooth_fix looks at it)
      o[i] = st[i].done(UB - UA + 1, rp);
      if (kmax - i < UA) o[i].x[0] = f_nan();
    }
  }

  // AGM fast path (agm.py:112-130, uniform weights_left / weights_right).  The running state of
  // UntilState<SEM_AGM> with (a) the shared tail -- the prefix accumulator of output i at q >= p0 + 8
  // is h_i + M_q, so l_q and r_q are lifted (one lg2 each) once per thread and position -- and (b)
  // the exp / log round trip between the prefix And and the pair And removed: when every l is
  // positive, log1p(y_i) IS the prefix mean of log1p(l) (y_i = expm1 of it), so the pair And reads the
  // accumulator directly.  Per candidate: one ex2 (geometric pair) and one lg2 (Or lift) -- 1 / cnt comes from an
  // eleven-entry table per group of four positions -- instead of about seven transcendentals.  Sums of logarithms
  // are carried in log2 units.
  PF_DEV static void phase2_agm(float* scr, const TC& c, float w0, float w1, V (&o)[8]) {
    const int f0 = c.p0 >> 2;
    constexpr int FMAX = LEXT / 4 - 1;
    const F4* L4 = reinterpret_cast<const F4*>(lplane(scr, 0));
    const F4* R4 = reinterpret_cast<const F4*>(rplane(scr, 0));
    const float invsw = fast_rcp(w0 + w1);
    const float w0p = w0 * invsw, w1p = w1 * invsw;  // weights_pair / their sum
    const int kmax = clip_kmax(c);
    float hneg[8], hlog[8], o0[8], o1[8];
    // one candidate of output i: prefix accumulator (pneg, plog), u = w0 / (sw cnt), right operand r (weighted by w1 / sw)
    // The Or lift of a candidate z <= 0 is lg2(1 - z): the factors 1 - z >= 1 of up to FOUR candidates are multiplied and one
    // logarithm taken of the product (|z| < 1e9 by the kernels' range assumption: the product stays below 1e36) -- 1.25 instead
    // of 2 MUFU operations per candidate on a walk whose XU pipe was 75 % busy.  A positive candidate makes the product +inf.
    float pr[8];
    auto cand = [&](int i, float pneg, float plog, float u, float rnegw, float rlogw, bool rpos) {
      const bool flag = plog == f_inf();  // some l <= 0: arithmetic prefix (agm.py:71)
      const float zA = fmaf(u, flag ? pneg : 0.f, rnegw);
      const float zG = fast_ex2(fmaf(u, plog, rlogw)) - 1.f;
      const float z = (flag || !rpos) ? zA : zG;
      const bool pos = z > 0.f;  // agm.py:86
      o0[i] += pos ? z : 0.f;
      pr[i] *= pos ? f_inf() : 1.f - z;
    };
    auto flush = [&](int i) {
      o1[i] += fast_lg2(pr[i]);
      pr[i] = 1.f;
    };

    {  // own chunk: triangular
      const F4 la = L4[pf4(f0)], lb = L4[pf4(f0 + 1)], ra = R4[pf4(f0)], rb = R4[pf4(f0 + 1)];
      const float lo[8] = {la.v[0], la.v[1], la.v[2], la.v[3], lb.v[0], lb.v[1], lb.v[2], lb.v[3]};
      const float ro[8] = {ra.v[0], ra.v[1], ra.v[2], ra.v[3], rb.v[0], rb.v[1], rb.v[2], rb.v[3]};
      float ln[8], ll[8], rn[8], rl[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        ln[q] = fminf(lo[q], 0.f);
        ll[q] = lo[q] <= 0.f ? f_inf() : fast_lg2(1.f + lo[q]);
        rn[q] = w1p * fminf(ro[q], 0.f);
        rl[q] = w1p * fast_lg2(1.f + ro[q]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float pn = 0.f, pl = 0.f;
        o0[i] = 0.f;
        o1[i] = 0.f;
        pr[i] = 1.f;
#pragma unroll
        for (int q = i; q < 8; ++q) {
          pn += ln[q];
          pl += ll[q];
          if (q - i >= UA && q - i <= UB && q <= kmax) cand(i, pn, pl, w0p * (1.f / (float)(q - i + 1)), rn[q], rl[q], ro[q] > 0.f);
          if (q - i == 3) flush(i);  // at most four factors per product
        }
        flush(i);
        hneg[i] = pn;
        hlog[i] = pl;
      }
    }
    float Mn = 0.f, Ml = 0.f;
    constexpr int KS0 = UA + 7 > 8 ? UA + 7 : 8;
    constexpr int KS = (KS0 + 3) & ~3;
    constexpr int KE = ((UB + 1) & ~3) > KS ? ((UB + 1) & ~3) : KS;  // body [KS, KE): every output takes every position
    // one 128-bit group of positions at offset 4 * g from p0; ALL: no per-candidate window test
    auto group = [&](int g, auto all) {
      const int fc = (f0 + g) < FMAX ? (f0 + g) : FMAX;
      const F4 ql = L4[pf4(fc)], qr = R4[pf4(fc)];
      // w0 / (sw cnt) for cnt = 4 g + (k - i) + 1, k - i in [-7, 3]: eleven reciprocals serve the group's 32 candidates
      float ik[11];
#pragma unroll
      for (int j = 0; j < 11; ++j) ik[j] = w0p * fast_rcp((float)(4 * g + j - 6));
      // (a variant of the group without the exponential for groups behind some l <= 0 was measured: cfg5-agm 6.39 -> 7.33 ms --
      // the lanes of a warp cross that point in different groups and then run both variants)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int kq = 4 * g + k;
        if (kq > kmax) break;
        const float l = ql.v[k], r = qr.v[k];
        Mn += fminf(l, 0.f);
        Ml += l <= 0.f ? f_inf() : fast_lg2(1.f + l);
        const float rn = w1p * fminf(r, 0.f), rl = w1p * fast_lg2(1.f + r);
        const bool rpos = r > 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (decltype(all)::value || (kq - i >= UA && kq - i <= UB)) cand(i, hneg[i] + Mn, hlog[i] + Ml, ik[k - i + 7], rn, rl, rpos);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) flush(i);  // the group's (at most four) factors of every output
    };
#pragma unroll
    for (int g = 2; g < KS / 4; ++g) group(g, BoolC<false>{});
#pragma unroll 1
    for (int g = KS / 4; g < KE / 4; ++g) group(g, BoolC<true>{});
#pragma unroll
    for (int g = KE / 4; g <= (UB + 7) / 4; ++g) group(g, BoolC<false>{});
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int last = kmax - i < UB ? kmax - i : UB;  // clipped candidate count (agm.py:25-26)
      const float invn = fast_rcp((float)(last - UA + 1));
      o[i].x[0] = last < UA ? f_nan() : ((o1[i] == f_inf()) ? o0[i] * invn : 1.f - fast_ex2(o1[i] * invn));
    }
  }

  // D-GMSR fast path (dgmsr.py:88-118 with p = 1, unit weights_pair, uniform weights_left / weights_right).  Same shared
  // tail as the AGM path on the prefix accumulator (sum of l^2 over l <= 0; sum of 2 lg2 l, or +inf once some l <= 0), and
  // the chain  y = And(prefix), z = And(y, r), U = Or(z)  written out per case instead of through the generic
  // lift / fin chain (thirteen transcendentals per candidate):
  //   prefix holds some l <= 0:  y = sqrt(eps) - sqrt(eps + a0 / cnt) <= 0,  z = sqrt(eps) - sqrt(eps + (y^2 + min(r, 0)^2) / 2)
  //                              -- two square roots, and one lg2 for the Or lift of z < 0;
  //   every l > 0, r <= 0:       z = sqrt(eps) - sqrt(eps + r^2 / 2) depends on the position only (y > 0 carries no weight);
  //   every l > 0, r > 0:        y = 2^(logaddexp2(cnt lg2 eps, a1) / (2 cnt)) - sqrt(eps),  z = (eps^2 + (y r)^2)^(1/4) - sqrt(eps)
  //                              (the two-operand form of dgmsr2_unit: two square roots, no logarithm); z > 0 joins the Or squared.
  // A group of positions behind some l <= 0 of the shared tail holds flagged prefixes only and takes the first case alone.
  // Assumes |l|, |r| < 1e9 (y r is squared in fp32).
  PF_DEV static void phase2_dgmsr(float* scr, const TC& c, const RP& rp, V (&o)[8]) {
    const int f0 = c.p0 >> 2;
    constexpr int FMAX = LEXT / 4 - 1;
    const F4* L4 = reinterpret_cast<const F4*>(lplane(scr, 0));
    const F4* R4 = reinterpret_cast<const F4*>(rplane(scr, 0));
    const int kmax = clip_kmax(c);
    const float eps = rp.eps_p, se = rp.sqrt_eps, lge = rp.ln_eps * kLog2e;  // (p = 1: eps_p = eps)
    float hn[8], hl[8], o0[8], o1[8];
    // the Or accumulator of output i takes candidate z (dgmsr.py:110-119: Or = -And(-z))
    auto take = [&](int i, float z) {
      const bool nn = z >= 0.f;
      o0[i] += nn ? z * z : 0.f;
      o1[i] += nn ? f_inf() : 2.f * fast_lg2(-z);
    };
    // candidate with a flagged prefix: a0 = sum of l^2 over the l <= 0 of the prefix, r2n = min(r, 0)^2
    auto cand_flag = [&](int i, float a0, float invk, float r2n) {
      const float y = se - fast_sqrt(fmaf(a0, invk, eps));
      take(i, se - fast_sqrt(fmaf(0.5f, fmaf(y, y, r2n), eps)));
    };
    // general candidate; zr: the value of an all-positive prefix paired with this r <= 0
    // (the three cases are BRANCHES: behind a few positions almost every prefix holds some l <= 0 and a whole warp takes the
    // two-square-root case -- with N(0, 1) signals the general case runs for about a third of the candidates of a warp;
    // PYSTL_DGMSR_SELECT=1 evaluates every case and selects, which is what this walk did before)
    auto cand = [&](int i, float a0, float a1, float invk, float kf, float r, float r2n, float zr) {
      const bool flag = a1 == f_inf();
#if PYSTL_DGMSR_SELECT == 0
      if (flag) {
        cand_flag(i, a0, invk, r2n);
        return;
      }
      if (!(r > 0.f)) {
        take(i, zr);
        return;
      }
#elif PYSTL_DGMSR_SELECT == 2
      if (__all_sync(__activemask(), flag)) {  // warp-uniform: no divergence, one vote per candidate
        cand_flag(i, a0, invk, r2n);
        return;
      }
#endif
      const float yf = se - fast_sqrt(fmaf(a0, invk, eps));
      const float zf = se - fast_sqrt(fmaf(0.5f, fmaf(yf, yf, r2n), eps));
      const float u = kf * lge;
      const float lae = fmaxf(u, a1) + fast_lg2(1.f + fast_ex2(-fabsf(u - a1)));
      const float yg = fast_ex2(0.5f * invk * lae) - se;
      const float q = yg * r;
      const float zg = fast_sqrt(fast_sqrt(fmaf(q, q, eps * eps))) - se;
      take(i, flag ? zf : (r > 0.f ? zg : zr));
    };
    {  // own chunk: triangular
      const F4 la = L4[pf4(f0)], lb = L4[pf4(f0 + 1)], ra = R4[pf4(f0)], rb = R4[pf4(f0 + 1)];
      const float lo[8] = {la.v[0], la.v[1], la.v[2], la.v[3], lb.v[0], lb.v[1], lb.v[2], lb.v[3]};
      const float ro[8] = {ra.v[0], ra.v[1], ra.v[2], ra.v[3], rb.v[0], rb.v[1], rb.v[2], rb.v[3]};
      float l2[8], ll[8], r2[8], zr[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float ln = fminf(lo[q], 0.f), rn = fminf(ro[q], 0.f);
        l2[q] = ln * ln;
        ll[q] = lo[q] <= 0.f ? f_inf() : 2.f * fast_lg2(lo[q]);
        r2[q] = rn * rn;
        zr[q] = se - fast_sqrt(fmaf(0.5f, r2[q], eps));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float pn = 0.f, pl = 0.f;
        o0[i] = 0.f;
        o1[i] = 0.f;
#pragma unroll
        for (int q = i; q < 8; ++q) {
          pn += l2[q];
          pl += ll[q];
          if (q - i >= UA && q - i <= UB && q <= kmax) cand(i, pn, pl, 1.f / (float)(q - i + 1), (float)(q - i + 1), ro[q], r2[q], zr[q]);
        }
        hn[i] = pn;
        hl[i] = pl;
      }
    }
    float Mn = 0.f, Ml = 0.f;
    constexpr int KS0 = UA + 7 > 8 ? UA + 7 : 8;
    constexpr int KS = (KS0 + 3) & ~3;
    constexpr int KE = ((UB + 1) & ~3) > KS ? ((UB + 1) & ~3) : KS;
    auto group = [&](int g, auto all) {
      const int fc = (f0 + g) < FMAX ? (f0 + g) : FMAX;
      const F4 ql = L4[pf4(fc)], qr = R4[pf4(fc)];
      float ik[11];  // 1 / cnt for cnt = 4 g + (k - i) + 1, k - i in [-7, 3]
#pragma unroll
      for (int j = 0; j < 11; ++j) ik[j] = fast_rcp((float)(4 * g + j - 6));
#if PYSTL_DGMSR_SELECT == 2
      if (__all_sync(__activemask(), Ml == f_inf())) {
#else
      if (Ml == f_inf()) {  // some l <= 0 in the shared tail already: every candidate of the group is flagged
#endif
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int kq = 4 * g + k;
          if (kq > kmax) break;
          const float ln = fminf(ql.v[k], 0.f), rn = fminf(qr.v[k], 0.f);
          Mn = fmaf(ln, ln, Mn);
          const float r2n = rn * rn;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (decltype(all)::value || (kq - i >= UA && kq - i <= UB)) cand_flag(i, hn[i] + Mn, ik[k - i + 7], r2n);
        }
        return;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int kq = 4 * g + k;
        if (kq > kmax) break;
        const float l = ql.v[k], r = qr.v[k];
        const float ln = fminf(l, 0.f), rn = fminf(r, 0.f);
        Mn = fmaf(ln, ln, Mn);
        Ml += l <= 0.f ? f_inf() : 2.f * fast_lg2(l);
        const float r2n = rn * rn;
        const float zr = se - fast_sqrt(fmaf(0.5f, r2n, eps));
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (decltype(all)::value || (kq - i >= UA && kq - i <= UB))
            cand(i, hn[i] + Mn, hl[i] + Ml, ik[k - i + 7], (float)(kq - i + 1), r, r2n, zr);
      }
    };
#pragma unroll
    for (int g = 2; g < KS / 4; ++g) group(g, BoolC<false>{});
#pragma unroll 1
    for (int g = KS / 4; g < KE / 4; ++g) group(g, BoolC<true>{});
#pragma unroll
    for (int g = KE / 4; g <= (UB + 7) / 4; ++g) group(g, BoolC<false>{});
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int last = kmax - i < UB ? kmax - i : UB;  // clipped candidate count (dgmsr.py:13-25)
      const float nf = (float)(last - UA + 1);
      const float invn = fast_rcp(nf);
      // Or = -And(-z): some z >= 0 -> arithmetic branch over those; else the geometric branch over -z > 0
      const float ha = se - fast_sqrt(fmaf(o0[i], invn, eps));
      const float u = nf * lge;
      const float lae = fmaxf(u, o1[i]) + fast_lg2(1.f + fast_ex2(-fabsf(u - o1[i])));
      const float hg = fast_ex2(0.5f * invn * lae) - se;
      o[i].x[0] = last < UA ? f_nan() : -((o1[i] == f_inf()) ? ha : hg);
    }
  }

  PF_DEV static void phase2(float* scr, const TC& c, const RP& rp, float w0, float w1, V (&o)[8], UntilState<SEM> (&st)[8]) {
    const int p0 = c.p0;
    if constexpr (SEM == SEM_DGMSR && UA <= 64) {
      if (PYSTL_P(rp) == 1 && w0 == 1.f && w1 == 1.f) {
        phase2_dgmsr(scr, c, rp, o);
        return;
      }
    }
    if constexpr (SEM == SEM_AGM && UA <= 64) {
      if (cum_fast_ok(c)) {
        phase2_agm(scr, c, w0, w1, o);
        return;
      }
    }
    if constexpr (SEM == SEM_CUMULATIVE && UA <= 64) {
      if (cum_fast_ok(c)) {
        phase2_cum(scr, c, o);
        return;
      }
    }
    if constexpr (SEM == SEM_SMOOTH && UA <= 64) {
      if (cum_fast_ok(c)) {
        phase2_smooth(scr, c, rp, o, st);
        return;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) st[i].init();
    int q_hi = p0 + 7 + UB;
    if (q_hi > c.nlive - 1) q_hi = c.nlive - 1;
#pragma unroll 1
    for (int f = p0 >> 2; f <= (q_hi >> 2); ++f) {
      F4 ql[VW], qr[VW];
#pragma unroll
      for (int w = 0; w < VW; ++w) {
        ql[w] = reinterpret_cast<const F4*>(lplane(scr, w))[pf4(f)];
        qr[w] = reinterpret_cast<const F4*>(rplane(scr, w))[pf4(f)];
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
          if (kk >= 0 && kk <= UB) st[i].step(l, r, kk >= UA, rp, w0, w1);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = ncand(c, p0 + i);
      if (n <= 0 || p0 + i >= c.nlive) {
#pragma unroll
        for (int w = 0; w < VW; ++w) o[i].x[w] = f_nan();
      } else {
        o[i] = st[i].done(n, rp);
      }
    }
  }

  // smooth only: outputs whose shifted sums lost their precision, recomputed with per-term scales
  PF_DEV static void smooth_fix(float* scr, const TC& c, const RP& rp, const UntilState<SEM_SMOOTH> (&st)[8], Vec<1> (&o)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a_i = st[i].acc;
      const int p = c.p0 + i;
      const int n = ncand(c, p);
      if (n <= 0 || p >= c.nlive || (a_i >= kSmoothTiny && a_i <= kSmoothUntilMax)) continue;
      o[i].x[0] = smooth_until_fix_one<UA>(p, n + UA - 1, rp.k, rp.invk, lplane(scr, 0), rplane(scr, 0));
    }
  }
};

// ------------------------------------------------------------------ weighted windows (agm / dgmsr)
// Always / Eventually with per-position weights (agm.py:18-35, dgmsr.py:13-25): weights[k] belongs
// to the k-th sample of the window and the list is truncated to the clipped window, so the
// reduction is a correlation -- O(W) per sample.  The unweighted lifted terms of every sample are
// staged once (both accumulator words are linear in the weight), each output then folds its
// window with its own weights.  The "an entry took the arithmetic branch" flag (+inf in word 1)
// does not depend on the weight, like np.any(values <= 0) in the reference.
template <int SEM, int MODE, int WA, int WB, int NT>
struct WeightedWinOp {
  using R = Red<SEM, MODE>;
  using A = typename R::A;
  static constexpr int LEXT = 8 * NT;
  static constexpr int FLOATS = 2 * LEXT;
  static constexpr int W = WB - WA + 1;

  PF_DEV static void phase1(float* scr, int tid, const A (&x)[8]) {
    float t[8];
    split_words<2>(x, 0, t);
    store8(scr, tid, t);
    split_words<2>(x, 1, t);
    store8(scr + LEXT, tid, t);
  }
  PF_DEV static void phase2(const float* scr, const TC& c, const RP& rp, const float* wts, Vec<1> (&o)[8]) {
    A acc[8];
    float sw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] = R::ident();
      sw[i] = 0.f;
    }
    const int p0 = c.p0;
    int q_hi = p0 + 7 + WB;
    if (q_hi > c.nlive - 1) q_hi = c.nlive - 1;
#pragma unroll 1
    for (int f = (p0 + WA) >> 2; f <= (q_hi >> 2); ++f) {
      const F4 t0 = reinterpret_cast<const F4*>(scr)[pf4(f)];
      const F4 t1 = reinterpret_cast<const F4*>(scr + LEXT)[pf4(f)];
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int q = 4 * f + k4;
        if (q > q_hi) break;
        const int d = q - p0 - WA;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = d - i;  // position inside output i's window
          if (k >= 0 && k < W) {
            const float w = wts[k];
            acc[i].x[0] += w * t0.v[k4];
            acc[i].x[1] += (t1.v[k4] == f_inf()) ? f_inf() : w * t1.v[k4];
            sw[i] += w;
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = sw[i] > 0.f ? R::fin(acc[i], sw[i], rp) : Vec<1>{{f_nan()}};
  }
};

// adjoint of a weighted window: d out_t / d x_q = in_factor(out_coeff(g_t, out_t, sw_t), x_q, w_{q-t-a}) is
// linear in the two-word coefficient, so g_src[q] = in_factor(sum_t w_{q-t-a} k_t, x_q, 1): the
// correlation of the staged coefficients with the weights, looking back over t in [q-b, q-a].
template <int SEM, int MODE, int WA, int WB, int NT>
struct WeightedWinAdj {
  using AD = Adj<SEM, MODE>;
  static constexpr int LEXT = 8 * NT;
  static constexpr int FLOATS = 2 * LEXT;
  static constexpr int W = WB - WA + 1;
  // sum of the weights of the clipped window of output p (agm.py:28-35: weights[:len])
  PF_DEV static float weight_sum(const TC& c, int p, const float* wts) {
    int n = c.nlive - 1 - p - WA + 1;
    if (n > W) n = W;
    float s = 0.f;
    for (int k = 0; k < n; ++k) s += wts[k];
    return s;
  }
  PF_DEV static void phase1(float* scr, const TC& c, const RP& rp, const float* wts, const float (&g)[8], const Vec<1> (&o)[8]) {
    Vec<2> k[8];
    const bool interior = c.p0 + 7 + WB <= c.nlive - 1;
    float sw_full = 0.f;
    if (interior)
      for (int j = 0; j < W; ++j) sw_full += wts[j];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float sw = interior ? sw_full : weight_sum(c, c.p0 + i, wts);
      bool lost;
      k[i] = (g[i] != 0.f && sw > 0.f) ? AD::out_coeff(g[i], o[i].x[0], sw, rp, lost) : Vec<2>{{0.f, 0.f}};
    }
    float t[8];
    split_words<2>(k, 0, t);
    store8(scr, c.tid, t);
    split_words<2>(k, 1, t);
    store8(scr + LEXT, c.tid, t);
  }
  template <bool NEG>
  PF_DEV static void phase2(const float* scr, const TC& c, const RP& rp, const float* wts, const Vec<1> (&v)[8], float (&gs)[8]) {
    Vec<2> acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = Vec<2>{{0.f, 0.f}};
    const int p0 = c.p0;
    int t_lo = p0 - WB;
    if (t_lo < 0) t_lo = 0;
    int t_hi = p0 + 7 - WA;
    if (t_hi > c.nlive - 1) t_hi = c.nlive - 1;
#pragma unroll 1
    for (int f = t_lo >> 2; f <= (t_hi >> 2); ++f) {
      const F4 k0 = reinterpret_cast<const F4*>(scr)[pf4(f)];
      const F4 k1 = reinterpret_cast<const F4*>(scr + LEXT)[pf4(f)];
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int t = 4 * f + k4;
        if (t < t_lo || t > t_hi) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = p0 + i - t - WA;  // position of sample p0 + i inside output t's window
          if (k >= 0 && k < W) {
            const float w = wts[k];
            acc[i].x[0] += w * k0.v[k4];
            acc[i].x[1] += w * k1.v[k4];
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x = NEG ? -v[i].x[0] : v[i].x[0];
      const float d = (c.p0 + i < c.nlive && x == x) ? AD::in_factor(acc[i], x, 1.f, rp) : 0.f;  // (x NaN: the reference raises there)
      gs[i] += NEG ? -d : d;
    }
  }
};

// Until with weights_left / weights_right (agm.py:90-130, dgmsr.py:88-118): weights_left[:i+1]
// over the prefix l[0..i], weights_pair on AND(prefix, r_i), weights_right over the candidates.
// wl / wr: null = uniform.  Same staging as UntilOp.
template <int SEM, int UA, int UB, int NT>
PF_DEV void until_weighted_phase2(float* scr, const TC& c, const RP& rp, float wp0, float wp1, const float* wl, const float* wr,
                                  Vec<1> (&o)[8]) {
  using U = UntilOp<SEM, UA, UB, NT>;
  using RA = Red<SEM, MODE_AND>;
  using RO = Red<SEM, MODE_OR>;
  Vec<2> pre[8], orr[8];
  float swl[8], swr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    pre[i] = RA::ident();
    orr[i] = RO::ident();
    swl[i] = swr[i] = 0.f;
  }
  const int p0 = c.p0;
  int q_hi = p0 + 7 + UB;
  if (q_hi > c.nlive - 1) q_hi = c.nlive - 1;
#pragma unroll 1
  for (int f = p0 >> 2; f <= (q_hi >> 2); ++f) {
    const F4 ql = reinterpret_cast<const F4*>(U::lplane(scr, 0))[pf4(f)];
    const F4 qr = reinterpret_cast<const F4*>(U::rplane(scr, 0))[pf4(f)];
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      const int q = 4 * f + k4;
      if (q > q_hi) break;
      const Vec<1> l{{ql.v[k4]}}, r{{qr.v[k4]}};
      const int d = q - p0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int kk = d - i;
        if (kk >= 0 && kk <= UB) {
          const float wlk = wl ? wl[kk] : 1.f;
          pre[i] = RA::comb(pre[i], RA::lift(l, rp, wlk));
          swl[i] += wlk;
          if (kk >= UA) {
            const Vec<1> y = RA::fin(pre[i], swl[i], rp);
            const Vec<2> pr = RA::comb(RA::lift(y, rp, wp0), RA::lift(r, rp, wp1));
            const float wrj = wr ? wr[kk - UA] : 1.f;
            orr[i] = RO::comb(orr[i], RO::lift(RA::fin(pr, wp0 + wp1, rp), rp, wrj));
            swr[i] += wrj;
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = U::ncand(c, p0 + i);
    o[i] = (n <= 0 || p0 + i >= c.nlive || !(swr[i] > 0.f)) ? Vec<1>{{f_nan()}} : RO::fin(orr[i], swr[i], rp);
  }
}

// ------------------------------------------------------------------ lift / finish
// the thread's operand samples -> accumulators; samples past the horizon become the identity
template <class R, int VW>
PF_DEV void lift8(const TC& c, const RP& rp, const Vec<VW> (&v)[8], typename R::A (&x)[8]) {
  if (own_all_live(c)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = R::lift(v[i], rp, 1.f);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = (c.p0 + i < c.nlive) ? R::lift(v[i], rp, 1.f) : R::ident();
  }
}

// accumulators -> values of the forward window [t + FA, min(t + FB, T - 1)] (api.py:49-54): the
// normaliser is the clipped sample count.  Empty windows (the reference raises) are masked at
// the root (start_slack), so any finite count will do for them.
template <class R, int FA, int FB, int VW>
PF_DEV void fin8(const TC& c, const RP& rp, const typename R::A (&acc)[8], Vec<VW> (&o)[8]) {
  if (c.p0 + 7 + FB <= c.nlive - 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = R::fin(acc[i], (float)(FB - FA + 1), rp);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int hi = c.p0 + i + FB;
      if (hi > c.nlive - 1) hi = c.nlive - 1;
      const int n = hi - (c.p0 + i + FA) + 1;
      o[i] = R::fin(acc[i], (float)(n > 0 ? n : 1), rp);
    }
  }
}

// smooth only: outputs whose block-shifted sum lost its precision to underflow are recomputed
// exactly from the staged operand (rare: needs a dynamic range above 2^100 inside one tile)
template <int MODE>
PF_COLD float smooth_fix_one(int lo, int hi, float k, float invk, const float* vplane) {
  using R = Red<SEM_SMOOTH, MODE>;
  if (lo > hi) return f_nan();
  RP rp{};
  rp.k = k;
  rp.invk = invk;
  float mx = -f_inf();
  for (int q = lo; q <= hi; ++q) mx = fmaxf(mx, R::expo(Vec<1>{{vplane[pidx(q)]}}, rp));
  float s = 0.f;
  for (int q = lo; q <= hi; ++q) s += fast_ex2(R::expo(Vec<1>{{vplane[pidx(q)]}}, rp) - mx);
  rp.shift = mx;
  return R::fin(Vec<1>{{s}}, 1.f, rp).x[0];
}
template <int MODE, int FA, int FB, int NT>
PF_DEV void smooth_fix8(const TC& c, const RP& rp, const float* vplane, const Vec<1> (&acc)[8], Vec<1> (&o)[8]) {
  float m = acc[0].x[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fminf(m, acc[i].x[0]);
  if (m >= kSmoothTiny) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (acc[i].x[0] >= kSmoothTiny || c.p0 + i >= c.nlive) continue;
    int hi = c.p0 + i + FB;
    if (hi > c.nlive - 1) hi = c.nlive - 1;
    if (hi > 8 * NT - 1) hi = 8 * NT - 1;
    o[i].x[0] = smooth_fix_one<MODE>(c.p0 + i + FA, hi, rp.k, rp.invk, vplane);
  }
}


// ------------------------------------------------------------------ AGM And / Or of TWO operands
// (agm.py:60-88; partials as Adj<SEM_AGM, MODE> gives them).  Same arithmetic as the generic
// lift / comb / fin / out_coeff / in_factor chain, written out for N = 2: two lg2, one ex2 and
// (with partials) two rcp, instead of the accumulator pairs and their selects.
template <int MODE, bool PARTIALS>
PF_DEV float agm2(float x0, float x1, float w0, float w1, float sw, float& d0, float& d1) {
  const float inv = fast_rcp(sw);
  const float u0 = w0 * inv, u1 = w1 * inv;
  if constexpr (MODE == MODE_AND) {
    const bool arith = x0 <= 0.f || x1 <= 0.f;  // agm.py:71
    const float E = fast_ex2(u0 * fast_lg2(1.f + x0) + u1 * fast_lg2(1.f + x1));
    const float o = arith ? u0 * fminf(x0, 0.f) + u1 * fminf(x1, 0.f) : E - 1.f;
    if constexpr (PARTIALS) {
      if (o <= 0.f) {
        d0 = x0 > 0.f ? 0.f : (x0 < 0.f ? u0 : 0.5f * u0);
        d1 = x1 > 0.f ? 0.f : (x1 < 0.f ? u1 : 0.5f * u1);
      } else {
        const float k = o + 1.f;
        d0 = x0 > 0.f ? u0 * k * fast_rcp(1.f + x0) : 0.f;
        d1 = x1 > 0.f ? u1 * k * fast_rcp(1.f + x1) : 0.f;
      }
    }
    return o;
  } else {
    const bool arith = x0 > 0.f || x1 > 0.f;  // agm.py:86
    const float E = fast_ex2(u0 * fast_lg2(1.f - x0) + u1 * fast_lg2(1.f - x1));
    const float o = arith ? u0 * fmaxf(x0, 0.f) + u1 * fmaxf(x1, 0.f) : 1.f - E;
    if constexpr (PARTIALS) {
      if (o >= 0.f) {
        d0 = x0 > 0.f ? u0 : (x0 < 0.f ? 0.f : 0.5f * u0);
        d1 = x1 > 0.f ? u1 : (x1 < 0.f ? 0.f : 0.5f * u1);
      } else {
        const float k = 1.f - o;
        d0 = x0 > 0.f ? 0.f : u0 * k * fast_rcp(1.f - x0);
        d1 = x1 > 0.f ? 0.f : u1 * k * fast_rcp(1.f - x1);
      }
    }
    return o;
  }
}

// ------------------------------------------------------------------ D-GMSR And / Or of TWO operands
// (_dgmsr.py:66-119 with p = 1 and unit weights, sw = 2), written out without the log domain:
//   some u <= 0:  h = sqrt(eps) - sqrt(eps + (sum_{u<=0} u^2) / 2),       dh/du_i = |u_i| / (2 sqrt(mp))  (u_i <= 0)
//   else:         m0 = sqrt(eps^2 + (u0 u1)^2),  h = sqrt(m0) - sqrt(eps),  dh/du_i = u0 u1 u_{1-i} / (2 m0^(3/2))
// with u = v (And) or -v (Or, which is -And(-v), _dgmsr.py:110-119).  Three square roots and two
// reciprocals instead of the thirteen transcendentals of the generic lift / fin / coefficient chain.
// The caller checks p == 1, w0 == w1 == 1 and |u0 u1| < 1e18 (the product is squared in fp32).
template <int MODE, bool PARTIALS>
PF_DEV float dgmsr2_unit(float x0, float x1, const RP& rp, float& d0, float& d1) {
  const float u0 = MODE == MODE_AND ? x0 : -x0, u1 = MODE == MODE_AND ? x1 : -x1;
  const bool arith = u0 <= 0.f || u1 <= 0.f;
  const float n0 = fminf(u0, 0.f), n1 = fminf(u1, 0.f);
  const float smp = fast_sqrt(fmaf(fmaf(n0, n0, n1 * n1), 0.5f, rp.eps_p));
  const float q = u0 * u1;
  const float m0 = fast_sqrt(fmaf(q, q, rp.eps_p * rp.eps_p));
  const float r0 = fast_sqrt(m0);
  const float h = arith ? rp.sqrt_eps - smp : r0 - rp.sqrt_eps;
  if constexpr (PARTIALS) {
    // (the coefficient branch of Adj<SEM_DGMSR>::out_coeff is chosen by h <= 0, like there)
    const bool neg = h <= 0.f;
    const float cn = 0.5f * fast_rcp(smp);
    const float cp = 0.5f * q * fast_rcp(m0 * r0);
    d0 = neg ? (u0 <= 0.f ? -n0 * cn : 0.f) : (u0 > 0.f ? cp * u1 : 0.f);
    d1 = neg ? (u1 <= 0.f ? -n1 * cn : 0.f) : (u1 > 0.f ? cp * u0 : 0.f);
  }
  return MODE == MODE_AND ? h : -h;
}

// ------------------------------------------------------------------ n-ary And / Or of one sample
template <int SEM, int MODE, int N>
PF_DEV Vec<Tr<SEM>::VW> nary1(const RP& rp0, const Vec<Tr<SEM>::VW> (&a)[N], const float (&w)[N], float sw) {
  using R = Red<SEM, MODE>;
  if constexpr (SEM == SEM_AGM && N == 2) {
    float d0, d1;
    return Vec<1>{{agm2<MODE, false>(a[0].x[0], a[1].x[0], w[0], w[1], sw, d0, d1)}};
  } else if constexpr (SEM == SEM_DGMSR && N == 2) {
    if (PYSTL_P(rp0) == 1 && w[0] == 1.f && w[1] == 1.f && fabsf(a[0].x[0] * a[1].x[0]) < 1e18f) {
      float d0, d1;
      return Vec<1>{{dgmsr2_unit<MODE, false>(a[0].x[0], a[1].x[0], rp0, d0, d1)}};
    }
    typename R::A acc = R::comb(R::lift(a[0], rp0, w[0]), R::lift(a[1], rp0, w[1]));
    return R::fin(acc, sw, rp0);
  } else if constexpr (SEM == SEM_SMOOTH && N == 2) {
    // extremum +- tau ln(1 + e^-|d|/tau): the max-shifted logsumexp of two terms (smooth.py:18-22)
    const float x0 = a[0].x[0], x1 = a[1].x[0];
    const float l = fast_lg2(1.f + fast_ex2(-fabsf(x0 - x1) * rp0.k));
    if constexpr (MODE == MODE_AND) return Vec<1>{{fmaf(l, -rp0.invk, fminf(x0, x1))}};
    else return Vec<1>{{fmaf(l, rp0.invk, fmaxf(x0, x1))}};
  } else if constexpr (SEM == SEM_SMOOTH) {
    // exact max-shifted logsumexp (smooth.py:18-22)
    float e[N];
    float m = -f_inf();
#pragma unroll
    for (int j = 0; j < N; ++j) {
      e[j] = R::expo(a[j], rp0);
      m = fmaxf(m, e[j]);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) s += fast_ex2(e[j] - m);
    RP rp = rp0;
    rp.shift = m;
    return R::fin(Vec<1>{{s}}, sw, rp);
  } else {
    typename R::A acc = R::lift(a[0], rp0, w[0]);
#pragma unroll
    for (int j = 1; j < N; ++j) acc = R::comb(acc, R::lift(a[j], rp0, w[j]));
    return R::fin(acc, sw, rp0);
  }
}

// ------------------------------------------------------------------ root -> HBM (+ fused mean)
// Writes the thread's outputs inside [p_lo, p_hi); NaN where the reference raises.  Returns the
// sum of the valid outputs it wrote (for the fused mean).
template <int SEM, bool NEG>
PF_DEV float store_root(const TC& c, long long T, const Vec<Tr<SEM>::VW> (&v)[8], float* __restrict__ out) {
  constexpr int VW = Tr<SEM>::VW;
  Vec<VW> o[8];
  float part = 0.f;
  if (c.p0 + 8 <= c.p_lo || c.p0 >= c.p_hi) return 0.f;  // halo thread: nothing of it is written
  float* dst = out + (c.b * T + c.origin + c.p0) * (long long)VW;
  const bool inside = c.p0 >= c.p_lo && c.p0 + 8 <= c.p_hi;
  if (inside && c.p0 + 7 <= c.last_ok) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = NEG ? negate<SEM>(v[i]) : v[i];
#pragma unroll
      for (int w = 0; w < VW; ++w) part += o[i].x[w];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = NEG ? negate<SEM>(v[i]) : v[i];
      const int p = c.p0 + i;
      if (p > c.last_ok) {
#pragma unroll
        for (int w = 0; w < VW; ++w) o[i].x[w] = f_nan();
      } else if (p >= c.p_lo && p < c.p_hi) {
#pragma unroll
        for (int w = 0; w < VW; ++w) part += o[i].x[w];
      }
    }
  }
  if (inside && ((u64)dst & 15ull) == 0ull) {
    F4* d4 = reinterpret_cast<F4*>(dst);
    if constexpr (VW == 1) {
      d4[0] = F4{{o[0].x[0], o[1].x[0], o[2].x[0], o[3].x[0]}};
      d4[1] = F4{{o[4].x[0], o[5].x[0], o[6].x[0], o[7].x[0]}};
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) d4[k] = F4{{o[2 * k].x[0], o[2 * k].x[1], o[2 * k + 1].x[0], o[2 * k + 1].x[1]}};
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (c.p0 + i >= c.p_lo && c.p0 + i < c.p_hi) {
#pragma unroll
        for (int w = 0; w < VW; ++w) dst[i * VW + w] = o[i].x[w];
      }
  }
  return part;
}

// block sum of `part` -> one atomicAdd(double) per CTA.  red: NT/32 floats.  All threads call it.
template <int NT>
PF_DEV void block_sum_to(const TC& c, float part, double* __restrict__ acc, float* red) {
  float s = part;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += simt::shfl_xor(s, d);
  if ((c.tid & 31) == 0) red[c.tid >> 5] = s;
  simt::sync();
  if (c.tid == 0) {
    double tot = 0.0;
    for (int w = 0; w < NT / 32; ++w) tot += (double)red[w];
    simt::atomic_add(acc, tot);
  }
}

// ====================================================================== backward pass
// The generated backward kernel first runs the forward program (values stay in registers: nothing
// was saved by the forward launch), then walks it in reverse.  Adjoints are register arrays
// g<slot>[8]; the adjoint of a window is a backward-looking window SUM (same WinOp, offsets
// [-b, -a]) of a per-output coefficient, times a per-input factor (tile_bwd.h: Adj<SEM, MODE>).

// clipped sample count of the forward window of output p (api.py:49-54)
template <int FA, int FB>
PF_DEV float win_count(const TC& c, int p) {
  int hi = p + FB;
  if (hi > c.nlive - 1) hi = c.nlive - 1;
  const int n = hi - (p + FA) + 1;
  return (float)(n > 0 ? n : 0);
}

// cotangent of the root for the thread's samples: grad_out (or the constant `scale`) for the tile
// positions below p_hi whose output exists, zero elsewhere
template <bool NEG>
PF_DEV void root_cotangent(const TC& c, long long T, const float* __restrict__ grad_out, float scale, float (&g)[8]) {
  const float* src = grad_out ? grad_out + (c.b * T + c.origin + c.p0) : nullptr;
  int hi = c.p_hi;
  if (hi > c.last_ok + 1) hi = c.last_ok + 1;
  if (c.p0 >= hi) {  // right-halo thread: its outputs belong to the next tile
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = 0.f;
    return;
  }
  if (c.p0 + 8 <= hi && (!src || ((u64)src & 15ull) == 0ull)) {
    if (src) {
      const F4 a = reinterpret_cast<const F4*>(src)[0], b = reinterpret_cast<const F4*>(src)[1];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[i] = NEG ? -a.v[i] : a.v[i];
        g[4 + i] = NEG ? -b.v[i] : b.v[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = NEG ? -scale : scale;
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v = 0.f;
    if (c.p0 + i < hi) v = src ? src[i] : scale;
    g[i] = NEG ? -v : v;
  }
}

// ---- classic value-and-gradient (values carry their source, SEM_CLASSIC_SRC)
// root (value, source) -> out; same masking and fused mean as store_root
template <bool NEG>
PF_DEV float store_root_src(const TC& c, long long T, const Vec<2> (&v)[8], float* __restrict__ out) {
  Vec<1> o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i].x[0] = v[i].x[0];
  return store_root<SEM_CLASSIC, NEG>(c, T, o, out);
}
// The gradient of a min / max chain is one-hot (torch.min / max(dim) route to the first extremal
// index, torch/semantics.py:106,113): each output adds its cotangent to the
// plane of its source predicate at its source position.  Consecutive outputs of a thread often share
// the source (a window minimum serves every output whose window contains it): runs are merged
// before the shared-memory atomic.  planes: one 8*NT plane per predicate, zeroed, store8 layout.
template <bool NEG, int NT>
PF_DEV void classic_scatter(const TC& c, long long T, const float* __restrict__ grad_out, float scale, const Vec<2> (&v)[8],
                            float* planes) {
  float g[8];
  root_cotangent<false>(c, T, grad_out, scale, g);
  int cur = 0;
  float acc = 0.f;
  bool open = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s = f2i(v[i].x[1]);
    const bool on = g[i] != 0.f;
    if (open && (!on || s != cur)) {
      const bool neg = (cur < 0) != NEG;
      simt::atomic_add(&planes[((cur >> 24) & 0x7f) * (8 * NT) + pidx(cur & 0xffffff)], neg ? -acc : acc);
      open = false;
    }
    if (on) {
      if (!open) {
        cur = s;
        acc = 0.f;
        open = true;
      }
      acc += g[i];
    }
  }
  if (open) {
    const bool neg = (cur < 0) != NEG;
    simt::atomic_add(&planes[((cur >> 24) & 0x7f) * (8 * NT) + pidx(cur & 0xffffff)], neg ? -acc : acc);
  }
}

// Same routing, but straight into the GRADIENT TILE as it lies in HBM: image[pos][d] (8*NT rows of D
// floats, zeroed).  Each run adds sign * coeff[k] * cotangent to the channels of its source predicate;
// the finished image leaves through ONE TMA bulk store per CTA (store_grad_image) instead of
// per-thread row stores -- the value-and-gradient kernel of the classic semantics is bound by the
// LSU data pipe (ncu: 86 %), and the uncoalesced 64-byte row stores were a quarter of its wavefronts.
template <bool NEG, int NT, int D>
PF_DEV void classic_scatter_rows(const TC& c, const TileProgram& P, long long T, const float* __restrict__ grad_out, float scale,
                                 const Vec<2> (&v)[8], float* image) {
  float g[8];
  root_cotangent<false>(c, T, grad_out, scale, g);
  int cur = 0;
  float acc = 0.f;
  bool open = false;
  auto flush = [&]() {
    const bool neg = (cur < 0) != NEG;
    const float a = neg ? -acc : acc;
    const int q = (cur >> 24) & 0x7f;
    float* row = image + (cur & 0xffffff) * D;
    for (int k = 0; k < P.preds[q].n_terms; ++k) simt::atomic_add(&row[P.preds[q].ch[k]], P.preds[q].coeff[k] * a);
  };
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s = f2i(v[i].x[1]);
    const bool on = g[i] != 0.f;
    if (open && (!on || s != cur)) {
      flush();
      open = false;
    }
    if (on) {
      if (!open) {
        cur = s;
        acc = 0.f;
        open = true;
      }
      acc += g[i];
    }
  }
  if (open) flush();
}
// zero the image with conflict-free 128-bit stores (all threads)
template <int NT, int D>
PF_DEV void zero_grad_image(float* image, int tid) {
  F4* p = reinterpret_cast<F4*>(image);
#pragma unroll
  for (int k = 0; k < 2 * D; ++k) p[k * NT + tid] = F4{{0.f, 0.f, 0.f, 0.f}};
}
// rows [p_lo, p_hi) of the image -> grad_signal.  All threads call it (after a barrier).
template <int NT, int D>
PF_DEV void store_grad_image(const TC& c, long long T, const float* image, float* __restrict__ grad_signal) {
  if (c.p_hi <= c.p_lo) return;
  float* dst = grad_signal + ((c.b * T + c.origin + c.p_lo) * (long long)D);
  const float* src = image + c.p_lo * D;
  const int n = (c.p_hi - c.p_lo) * D;  // floats; p_lo, p_hi multiples of 8 except at the end of the trace
  if ((((u64)dst | (u64)src) & 15ull) == 0ull && (n & 3) == 0) {
    if (c.tid == 0) simt::bulk_store_and_wait(dst, src, (unsigned)n * 4u);
  } else {
    for (int j = c.tid; j < n; j += NT) dst[j] = src[j];
  }
}

// ---- window adjoint, phase A: per-output coefficients k[8] (KW words)
// smooth: d out_t / d x_q = xhat_q / S_t with the forward's block-shifted xhat and S: k_t = g_t / S_t.
// Outputs whose shifted sum lost its precision (S_t < 2^-100, recomputed exactly by the forward
// pass) are left out here and scattered exactly afterwards (lostmask).
// minimum that PROPAGATES a NaN (fminf drops it): SASS FMNMX.NAN
PF_DEV float fmin_nan(float a, float b) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
#else
  return (a != a || b != b) ? f_nan() : fminf(a, b);
#endif
}
template <int MODE, int FA, int FB>
PF_DEV unsigned smooth_adj_coeff(const TC& c, const float (&g)[8], const Vec<1> (&S)[8], Vec<1> (&k)[8]) {
  // (a window over an operand that is NaN where the reference raises -- an Until with a lower bound, near the end of the
  // trace -- has a NaN sum and a zero cotangent: the NaN must send the thread down the guarded path, 0 * rcp(NaN) is NaN)
  float m = S[0].x[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmin_nan(m, S[i].x[0]);
  if (m >= kSmoothTiny && c.p0 + 7 + FA <= c.nlive - 1) {  // every output has a window and a healthy sum
#pragma unroll
    for (int i = 0; i < 8; ++i) k[i].x[0] = g[i] * fast_rcp(S[i].x[0]);
    return 0u;
  }
  unsigned lost = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool has = g[i] != 0.f && c.p0 + i + FA <= c.nlive - 1;
    const bool ok = S[i].x[0] >= kSmoothTiny;
    k[i].x[0] = (has && ok) ? g[i] * fast_rcp(S[i].x[0]) : 0.f;
    if (has && !ok) lost |= 1u << i;
  }
  return lost;
}
template <int SEM, int MODE, int FA, int FB>
PF_DEV void gm_adj_coeff(const TC& c, const RP& rp, const float (&g)[8], const Vec<1> (&o)[8], Vec<2> (&k)[8]) {
  using AD = Adj<SEM, MODE>;
  if (c.p0 + 7 + FB <= c.nlive - 1) {  // no window of this thread is clipped
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      bool lost;
      k[i] = AD::out_coeff(g[i], o[i].x[0], (float)(FB - FA + 1), rp, lost);
      if (!(g[i] != 0.f)) k[i] = Vec<2>{{0.f, 0.f}};  // (a NaN output has a zero cotangent)
    }
  } else {  // clipped at the horizon (or a halo output at the end of the tile)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float n = win_count<FA, FB>(c, c.p0 + i);
      bool lost;
      k[i] = AD::out_coeff(g[i], o[i].x[0], n > 0.f ? n : 1.f, rp, lost);
      if (!(g[i] != 0.f) || !(n > 0.f)) k[i] = Vec<2>{{0.f, 0.f}};
    }
  }
}

// ---- window adjoint, phase B: g_src[q] += sign * in_factor(x_q) * (window sum of k)[q]
// NANSAFE: the operand can be NaN (where the reference raises: an Until with a lower bound or a weighted window near the end
// of the trace; codegen.cpp: may_nan).  Every output that reads such a sample has a zero cotangent, so its window sum w is 0
// there -- and 0 * NaN must not become the operand's adjoint.
template <bool NEG, bool NANSAFE>
PF_DEV void smooth_adj_apply(const Vec<1> (&w)[8], const Vec<1> (&xhat)[8], float (&gs)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float d = w[i].x[0] * xhat[i].x[0];  // xhat is 0 past the horizon
    if (NANSAFE) d = xhat[i].x[0] == xhat[i].x[0] ? d : 0.f;
    gs[i] += NEG ? -d : d;
  }
}
template <int SEM, int MODE, bool NEG, bool NANSAFE>
PF_DEV void gm_adj_apply(const TC& c, const RP& rp, const Vec<2> (&w)[8], const Vec<1> (&v)[8], float (&gs)[8]) {
  using AD = Adj<SEM, MODE>;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float x = NEG ? -v[i].x[0] : v[i].x[0];
    const bool ok = c.p0 + i < c.nlive && (!NANSAFE || x == x);
    const float d = ok ? AD::in_factor(w[i], x, 1.f, rp) : 0.f;
    gs[i] += NEG ? -d : d;
  }
}

// smooth: exact scatter of the outputs in `lostmask` (rare path, entered by the whole CTA when any
// thread has one).  vplane: staged operand values (after negation); fix: zeroed plane.
template <int MODE>
PF_COLD void smooth_adj_scatter_one(int lo, int hi, float gi, float oi, float k, const float* vplane, float* fix) {
  using AD = Adj<SEM_SMOOTH, MODE>;
  RP rp{};
  rp.k = k;
  for (int q = lo; q <= hi; ++q) simt::atomic_add(&fix[pidx(q)], gi * AD::pair_weight(oi, vplane[pidx(q)], rp));
}
template <int MODE, int FA, int FB, int NT>
PF_DEV void smooth_adj_scatter(const TC& c, const RP& rp, unsigned lostmask, const float (&g)[8], const Vec<1> (&o)[8],
                               const float* vplane, float* fix) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (!((lostmask >> i) & 1)) continue;
    const int p = c.p0 + i;
    int hi = p + FB;
    if (hi > c.nlive - 1) hi = c.nlive - 1;
    if (hi > 8 * NT - 1) hi = 8 * NT - 1;
    smooth_adj_scatter_one<MODE>(p + FA, hi, g[i], o[i].x[0], rp.k, vplane, fix);
  }
}

// ---- n-ary And / Or of one sample with its local partials d out / d a_j, computed together in the
// forward sweep so that the backward sweep only multiplies (and the operands need not stay live)
template <int SEM, int MODE, int N>
PF_DEV Vec<1> nary1_partials(const RP& rp, const Vec<1> (&a)[N], const float (&w)[N], float sw, float (&d)[N]) {
  using R = Red<SEM, MODE>;
  using AD = Adj<SEM, MODE>;
  if constexpr (SEM == SEM_AGM && N == 2) {
    return Vec<1>{{agm2<MODE, true>(a[0].x[0], a[1].x[0], w[0], w[1], sw, d[0], d[1])}};
  } else if constexpr (SEM == SEM_DGMSR && N == 2) {
    if (PYSTL_P(rp) == 1 && w[0] == 1.f && w[1] == 1.f && fabsf(a[0].x[0] * a[1].x[0]) < 1e18f)
      return Vec<1>{{dgmsr2_unit<MODE, true>(a[0].x[0], a[1].x[0], rp, d[0], d[1])}};
    const Vec<1> o = nary1<SEM, MODE, N>(rp, a, w, sw);
    bool lost;
    const Vec<AD::KW> k = AD::out_coeff(1.f, o.x[0], sw, rp, lost);
#pragma unroll
    for (int j = 0; j < N; ++j) d[j] = AD::in_factor(k, a[j].x[0], w[j], rp);
    return o;
  } else if constexpr (SEM == SEM_SMOOTH && N == 2) {
    // softmax / softmin weights of two terms: 1 / (1 + 2^-|d|) for the extremal one
    // (two operands at the identity -- both windows empty, the reference raises -- give inf - inf: the minimum
    // drops the NaN, so the partials are finite everywhere and the adjoint sweep needs no guard)
    const float x0 = a[0].x[0], x1 = a[1].x[0];
    const float t = fast_ex2(-fminf(fabsf(x0 - x1), 3.0e38f) * rp.k);
    const float s = 1.f + t;
    const float wmax = fast_rcp(s), wmin = t * wmax;
    const bool first = MODE == MODE_AND ? x0 <= x1 : x0 >= x1;  // operand 0 is the extremal one
    d[0] = first ? wmax : wmin;
    d[1] = first ? wmin : wmax;
    if constexpr (MODE == MODE_AND) return Vec<1>{{fmaf(fast_lg2(s), -rp.invk, fminf(x0, x1))}};
    else return Vec<1>{{fmaf(fast_lg2(s), rp.invk, fmaxf(x0, x1))}};
  } else {
    const Vec<1> o = nary1<SEM, MODE, N>(rp, a, w, sw);
    if constexpr (SEM == SEM_SMOOTH) {
#pragma unroll
      for (int j = 0; j < N; ++j) d[j] = AD::pair_weight(o.x[0], a[j].x[0], rp);
    } else {
      bool lost;
      const Vec<AD::KW> k = AD::out_coeff(1.f, o.x[0], sw, rp, lost);
#pragma unroll
      for (int j = 0; j < N; ++j) d[j] = AD::in_factor(k, a[j].x[0], w[j], rp);
    }
    return o;
  }
}

PF_DEV void plane_zero8(float* plane, int tid) {
  const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  store8(plane, tid, z);
}

// ---- smooth until, backward.  U = tau ln A, A = sum_{i>=UA} t_i, t_i = 1 / (S_i + rho_i),
// S_i = sum_{k<=i} lambda_k, lambda = e^{-l/tau}, rho = e^{-r/tau}, hence
//   dU/dr_i = t_i^2 rho_i / A,   dU/dl_k = lambda_k (sum_{i >= max(k, UA)} t_i^2) / A.
// Everything is carried in the log2 domain (c_i = log2 t_i in a per-thread scratch array), so the
// weights are exact for any temperature: every exponent below is <= log2(b + 1).
PF_DEV float lse2(float a, float b) {  // log2(2^a + 2^b)
  const float m = fmaxf(a, b);
  return m + fast_lg2(fast_ex2(a - m) + fast_ex2(b - m));
}
template <int UA, int UB>
PF_COLD void smooth_until_scatter_one(int p, int last, float gi, float k, const float* lpl, const float* rpl, float* gl, float* gr) {
  float cbuf[UB + 1];
  float LS = -f_inf(), LA = -f_inf();
  for (int kk = 0; kk <= last; ++kk) {
    const int q = pidx(p + kk);
    const float zl = -lpl[q] * k;
    LS = kk == 0 ? zl : lse2(LS, zl);
    if (kk >= UA) {
      const float ci = -lse2(LS, -rpl[q] * k);
      cbuf[kk] = ci;
      LA = kk == UA ? ci : lse2(LA, ci);
    }
  }
  float LD = -f_inf();
  for (int kk = last; kk >= 0; --kk) {
    const int q = pidx(p + kk);
    if (kk >= UA) {
      const float ci = cbuf[kk];
      LD = kk == last ? 2.f * ci : lse2(LD, 2.f * ci);
      simt::atomic_add(&gr[q], gi * fast_ex2(2.f * ci - rpl[q] * k - LA));
    }
    simt::atomic_add(&gl[q], gi * fast_ex2(LD - lpl[q] * k - LA));
  }
}
template <int UA, int UB, int NT>
PF_DEV void smooth_until_scatter(const TC& c, const RP& rp, unsigned mask, const float (&g)[8], const float* lpl, const float* rpl, float* gl,
                                 float* gr) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (!((mask >> i) & 1u)) continue;
    const int p = c.p0 + i;
    int last = c.nlive - 1 - p;
    if (last > UB) last = UB;
    if (g[i] == 0.f || last < UA) continue;
    smooth_until_scatter_one<UA, UB>(p, last, g[i], rp.k, lpl, rpl, gl, gr);
  }
}
// ---- smooth until, backward, FAST PATH: the adjoint of UntilOp::phase2_smooth in the block-shifted linear domain.
// With L_k = 2^(-k l_k - shift), R_q = 2^(-k r_q - shift), S_t(q) = sum_{k = t..q} L_k, T_t(q) = 1 / (S_t(q) + R_q) and
// A_t = sum_q T_t(q) (what the forward pass kept), c_t = g_t / A_t:
//     grad_r[q] = R_q sum_t c_t T_t(q)^2,        grad_l[k] = L_k sum_t c_t Z_t(k),   Z_t(k) = sum_{q >= max(k, t + UA)} T_t(q)^2.
// The 8 outputs of a thread walk their candidates together, like the forward pass: past the own chunk S_i(q) = hS_i + M_q, so
// a candidate costs one add, one reciprocal and three multiply-adds.  The walk runs DOWNWARDS in q, so that the suffix sums
// Z_i accumulate without any subtraction (sums of positive terms only); M_q comes from the staged chunk aggregates
// (cs[]: running sums of whole chunks, built upwards first) plus an in-register prefix of the chunk at hand.  Per position
// the thread hands TWO numbers to the position's owner (the sum over its outputs of Z_i -- a running sum of the other
// number -- and of c_i T^2) instead of two shared-memory atomics per candidate; the own chunk is triangular and stays in registers.  L_k and R_q are exponentiated once per position by their owner and
// staged (la / rh / ag), and applied by the owner at the very end.
// Outputs whose A_t left the healthy range (the forward pass recomputed them exactly, smooth_fix) get c = 0 here and go
// through the exact log-domain scatter above (lost mask; block-uniform rare path).
template <int NT>
PF_DEV void smooth_until_adj_stage(const TC& c, const RP& rp, const float* lpl, const float* rpl, float* la, float* rh, float* ag) {
  float l[8], r[8], lam[8], rho[8];
  load8(lpl, c.tid, l);
  load8(rpl, c.tid, r);
  const float nk = -rp.k, ns = -rp.shift;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool ok = c.p0 + i < c.nlive;
    lam[i] = ok ? fast_ex2(fmaf(l[i], nk, ns)) : 0.f;
    rho[i] = ok ? fast_ex2(fmaf(r[i], nk, ns)) : 0.f;
    s += lam[i];
  }
  store8(la, c.tid, lam);
  store8(rh, c.tid, rho);
  ag[c.tid] = s;
}
// c_i = g_i / A_i for the outputs with a cotangent and a healthy sum; returns the mask of the outputs for the exact path
template <int UA, int UB>
PF_DEV unsigned smooth_until_adj_coeff(const TC& c, const float (&g)[8], const float (&A)[8], float (&ci)[8]) {
  unsigned lost = 0;
  const int kmax = c.nlive - 1 - c.p0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool has = g[i] != 0.f && kmax - i >= UA;
    const bool ok = A[i] >= kSmoothTiny && A[i] <= kSmoothUntilMax;
    ci[i] = (has && ok) ? g[i] * fast_rcp(A[i]) : 0.f;
    if (has && !ok) lost |= 1u << i;
  }
  return lost;
}
// All threads of the CTA call it (it contains barriers).  The chunks are visited in LOCKSTEP: in step cc every thread
// finishes the positions of chunk tid + cc and adds its two partial sums per position to that chunk's slots of gl / gr --
// one writer per chunk and step, so a plain 128-bit read-modify-write instead of atomics (a shared-memory float atomic
// is a compare-and-swap loop) and ONE barrier per step; the owner collects its slots once, after the last step.
template <int UA, int UB, int NT>
PF_DEV void smooth_until_adj_walk(const TC& c, bool live, const float* la, const float* rh, const float* ag, float* gl, float* gr,
                                  const float (&ci)[8], float (&gl_own)[8], float (&gr_own)[8]) {
  static_assert(UA >= 0 && UB >= UA, "bounded until");
  constexpr int CMAX = (UB + 7) / 8;  // last chunk (relative to the own one) that holds a candidate
  const int kmax = c.interior ? 0x3fffffff : c.nlive - 1 - c.p0;  // windows are clipped at the horizon on the last tile
  bool active = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) active = active || ci[i] != 0.f;
  active = active && live;
  float lam[8], rho[8], hS[8], Z[8];
  load8(la, c.tid, lam);
  load8(rh, c.tid, rho);
  {
    float s = 0.f;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      s += lam[i];
      hS[i] = ci[i] != 0.f ? s : 1.f;  // outputs without a cotangent: keep their reciprocals finite (0 * inf)
      Z[i] = 0.f;
    }
  }
  // chunks past the tile (interior tiles: only outputs with a zero cotangent reach them) or past the horizon are skipped
  int cend = CMAX;
  if (cend > NT - 1 - c.tid) cend = NT - 1 - c.tid;
  if (cend > (kmax >> 3)) cend = kmax >> 3;
  if (!active) cend = 0;
  float cs[CMAX + 1];  // cs[j] = sum of the aggregates of chunks tid + 1 .. tid + j
  cs[0] = 0.f;
#pragma unroll 1
  for (int cc = 1; cc <= cend; ++cc) cs[cc] = cs[cc - 1] + ag[c.tid + cc];
  float zs = 0.f;  // sum over the thread's outputs of Z_i: the running sum of the per-position sums
  // the positions of chunk tid + cc, walked downwards.  ALL: every (position, output) pair is a candidate;
  // CLIP: the chunk straddles the horizon
  auto chunk = [&](int cc, auto all, auto clip, float (&zc)[8], float (&rc)[8]) {
    float L[8], R[8];
    load8(la, c.tid + cc, L);
    load8(rh, c.tid + cc, R);
#pragma unroll
    for (int k = 1; k < 8; ++k) L[k] += L[k - 1];
    const float base = cs[cc - 1];
#pragma unroll
    for (int k = 7; k >= 0; --k) {
      float rs = 0.f;
      if (!decltype(clip)::value || 8 * cc + k <= kmax) {
        const float G = base + L[k] + R[k];
        if constexpr (decltype(all)::value) {
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            float t0, t1;
            rcp2(G + hS[i], G + hS[i + 1], t0, t1);
            const float v0 = t0 * ci[i], v1 = t1 * ci[i + 1];
            Z[i] = fmaf(v0, t0, Z[i]);
            Z[i + 1] = fmaf(v1, t1, Z[i + 1]);
            rs = fmaf(v0, t0, rs);
            rs = fmaf(v1, t1, rs);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (8 * cc + k - i >= UA && 8 * cc + k - i <= UB) {
              const float t = fast_rcp(G + hS[i]);
              const float v = t * ci[i];
              Z[i] = fmaf(v, t, Z[i]);
              rs = fmaf(v, t, rs);
            }
        }
      }
      zs += rs;
      zc[k] = zs;
      rc[k] = rs;
    }
  };
  auto step = [&](int cc, auto all) {
    float zc[8], rc[8];
    if (cc <= cend) {
      if (8 * cc + 7 <= kmax) chunk(cc, all, BoolC<false>{}, zc, rc);
      else chunk(cc, all, BoolC<true>{}, zc, rc);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) zc[k] = rc[k] = 0.f;
    }
    if (cc == CMAX) {  // the first step initialises every slot: its writer's numbers, or zeros where there is no writer
      if (c.tid + cc < NT) {
        store8(gl, c.tid + cc, zc);
        store8(gr, c.tid + cc, rc);
      }
      if (c.tid < CMAX) {
        plane_zero8(gl, c.tid);
        plane_zero8(gr, c.tid);
      }
    } else if (c.tid + cc < NT && cc <= cend) {  // one writer per chunk and step: a plain read-modify-write
      float a[8], b[8];
      load8(gl, c.tid + cc, a);
      load8(gr, c.tid + cc, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a[k] += zc[k];
        b[k] += rc[k];
      }
      store8(gl, c.tid + cc, a);
      store8(gr, c.tid + cc, b);
    }
    simt::sync();
  };
  // static classification of the chunks: cc is full when 8 cc - 7 >= UA and 8 cc + 7 <= UB
  constexpr int CF0 = (UA + 7 + 7) / 8 > 1 ? (UA + 7 + 7) / 8 : 1;  // first full chunk
  constexpr int CF1 = UB >= 15 ? (UB - 7) / 8 : 0;                   // last full chunk (0: none)
  if constexpr (CF1 >= CF0) {
#pragma unroll
    for (int cc = CMAX; cc > CF1; --cc) step(cc, BoolC<false>{});
#pragma unroll 1
    for (int cc = CF1; cc >= CF0; --cc) step(cc, BoolC<true>{});
#pragma unroll
    for (int cc = CF0 - 1; cc >= 1; --cc) step(cc, BoolC<false>{});
  } else {
#pragma unroll
    for (int cc = CMAX; cc >= 1; --cc) step(cc, BoolC<false>{});
  }
  if (live) {  // what the threads before this one handed over
    float a[8], b[8];
    load8(gl, c.tid, a);
    load8(gr, c.tid, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      gl_own[k] += a[k];
      gr_own[k] += b[k];
    }
  }
  if (!active) return;
  // own chunk: triangular, in registers
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float tmp[8];
    float S = ci[i] != 0.f ? 0.f : 1.f;
#pragma unroll
    for (int q = i; q < 8; ++q) {
      S += lam[q];
      tmp[q] = 0.f;
      if (q - i >= UA && q - i <= UB && q <= kmax) {
        const float t = fast_rcp(S + rho[q]);
        const float e = (t * ci[i]) * t;
        tmp[q] = e;
        gr_own[q] += e;
      }
    }
    float run = Z[i];
#pragma unroll
    for (int q = 7; q >= i; --q) {
      run += tmp[q];
      gl_own[q] += run;
    }
  }
}
// owner: apply L_k / R_q, route through the operands' negations
template <bool NL, bool NR>
PF_DEV void smooth_until_adj_finish(const float* la, const float* rh, int tid, const float (&gl_own)[8], const float (&gr_own)[8],
                                    float (&gsl)[8], float (&gsr)[8]) {
  float lam[8], rho[8];
  load8(la, tid, lam);
  load8(rh, tid, rho);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float dl = lam[i] * gl_own[i], dr = rho[i] * gr_own[i];
    gsl[i] += NL ? -dl : dl;
    gsr[i] += NR ? -dr : dr;
  }
}
// ---- agm until, backward, FAST PATH (uniform weights_left / weights_right): the adjoint of UntilOp::phase2_agm on the
// same lockstep scheme as smooth_until_adj_walk.  With y_i the prefix And of l over cnt samples, z = And((y_i, r_q); w0, w1)
// and U = Or_q z (agm.py:112-130), the chain of Adj<SEM_AGM> partials collapses per candidate to
//     dz  = g dU/dz                       (U >= 0: [z > 0] / n, else (1 - U) / (n (1 - z)))
//     kp0 = [z <= 0] dz / sw,  kp1 = [z > 0] dz (1 + z) / sw                              (pair And, out_coeff)
//     r-route:  (r > 0 ? kp1 / (1 + r) : kp0 [r < 0]) w1          -- the factor in r is applied by the owner of q
//     l-route:  two-word suffix sums over the candidates  Z0 += w0 kp0 [y < 0] / cnt,  Z1 += w0 kp1 / cnt
//               (the (1 + y) of in_factor and of out_coeff cancel), applied by the owner of k as
//               l_k > 0 ? Z1 / (1 + l_k) : Z0 [l_k < 0]   (ties: one half, torch.minimum)
// so a candidate costs an ex2 and a rcp (1 / cnt comes from a 15-entry table per chunk) and about thirty selects / FMAs;
// a chunk behind some l <= 0 (every prefix that reaches it is arithmetic) takes a variant without the exponential.
template <int NT>
PF_DEV void agm_until_adj_stage(const TC& c, const float* lpl, float* agn, float* agl) {
  float l[8];
  load8(lpl, c.tid, l);
  float sn = 0.f, sl = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (c.p0 + i < c.nlive) {
      sn += fminf(l[i], 0.f);
      sl += l[i] <= 0.f ? f_inf() : fast_lg2(1.f + l[i]);
    }
  agn[c.tid] = sn;
  agl[c.tid] = sl;
}
// kg_i = g_i / n_i (U_i >= 0) or g_i (1 - U_i) / n_i: the output-level coefficient of the Or; upos: bit i = (U_i >= 0)
template <int UA, int UB>
PF_DEV unsigned agm_until_adj_coeff(const TC& c, const float (&g)[8], const Vec<1> (&U)[8], float (&kg)[8]) {
  unsigned upos = 0;
  const int kmax = c.nlive - 1 - c.p0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int last = kmax - i;
    if (last > UB) last = UB;
    const int n = last - UA + 1;
    const bool has = g[i] != 0.f && n > 0;
    const bool pos = U[i].x[0] >= 0.f;
    const float gn = g[i] * fast_rcp((float)(n > 0 ? n : 1));
    kg[i] = has ? (pos ? gn : gn * (1.f - U[i].x[0])) : 0.f;
    if (pos) upos |= 1u << i;
  }
  return upos;
}
template <int UA, int UB, int NT>
PF_DEV void agm_until_adj_walk(const TC& c, bool live, float w0, float w1, const float* lpl, const float* rpl, const float* agn,
                               const float* agl, float* g0p, float* g1p, float* grp, const float (&kg)[8], unsigned upos,
                               float (&z0_own)[8], float (&z1_own)[8], float (&r_own)[8]) {
  static_assert(UA >= 0 && UB >= UA, "bounded until");
  constexpr int CMAX = (UB + 7) / 8;
  const int kmax = c.interior ? 0x3fffffff : c.nlive - 1 - c.p0;
  const float invsw = fast_rcp(w0 + w1);
  const float w0p = w0 * invsw, w1p = w1 * invsw;  // (the r-route's own 1 / sw is applied by the owner)
  bool active = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) active = active || kg[i] != 0.f;
  active = active && live;
  // coefficient of the Or per output: kS for a candidate z > 0, kG / (1 - z) for z < 0 (zero when U >= 0: the arithmetic
  // Or ignores its negative entries), their mean on the tie z == 0 (Adj<SEM_AGM, MODE_OR>::in_factor)
  float kG[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) kG[i] = ((upos >> i) & 1u) ? 0.f : kg[i];
  float lo[8], ro[8], lnq[8], llq[8], hneg[8], hlog[8], Z0[8], Z1[8];
  load8(lpl, c.tid, lo);
  load8(rpl, c.tid, ro);
  {
    float sn = 0.f, sl = 0.f;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      const bool ok = c.p0 + i < c.nlive;
      lnq[i] = ok ? fminf(lo[i], 0.f) : 0.f;
      llq[i] = ok ? (lo[i] <= 0.f ? f_inf() : fast_lg2(1.f + lo[i])) : 0.f;
      sn += lnq[i];
      sl += llq[i];
      // outputs without a cotangent take the arithmetic branch everywhere: nothing of theirs can overflow
      hneg[i] = kg[i] != 0.f ? sn : 0.f;
      hlog[i] = kg[i] != 0.f ? sl : f_inf();
      Z0[i] = Z1[i] = 0.f;
    }
  }
  // one candidate of output i.  pn / pl: prefix accumulator (sum of min(l, 0); sum of lg2(1 + l) or +inf once some l <= 0),
  // up = w0 / (sw cnt), rnp = w1 min(r, 0) / sw, rlp = w1 lg2(1 + r) / sw, rpos = r > 0.
  //   -> d0 / d1: contributions to Z0 / Z1, kr: to the r-route (times sw)
  auto cand = [&](int i, float pn, float pl, float up, float rnp, float rlp, bool rpos, float& d0, float& d1, float& kr) {
    const bool flag = pl == f_inf();  // some l <= 0: arithmetic prefix (agm.py:71)
    const float zA = fmaf(up, flag ? pn : 0.f, rnp);
    const float E = fast_ex2(fmaf(up, pl, rlp));  // 1 + z of the geometric pair
    const float z = (flag || !rpos) ? zA : E - 1.f;
    const float rz = fast_rcp(1.f - fminf(z, 0.f));
    float dz = z > 0.f ? kg[i] : kG[i] * rz;
    dz = z == 0.f ? 0.5f * (kg[i] + kG[i]) : dz;
    const float tE = z > 0.f ? dz * E : 0.f;
    d0 = flag ? up * dz * (pn < 0.f ? 1.f : 0.5f) : 0.f;
    d1 = up * tE;
    kr = rpos ? tE : dz;
  };
  // the same when the prefix is known to hold some l <= 0 (pl = +inf): arithmetic pair, z <= 0, no exponential
  auto cand_arith = [&](int i, float pn, float up, float rnp, float& d0, float& dz) {
    const float z = fmaf(up, pn, rnp);
    const float rz = fast_rcp(1.f - z);
    dz = z < 0.f ? kG[i] * rz : 0.5f * (kg[i] + kG[i]);
    d0 = up * dz * (pn < 0.f ? 1.f : 0.5f);
  };
  int cend = CMAX;
  if (cend > NT - 1 - c.tid) cend = NT - 1 - c.tid;
  if (cend > (kmax >> 3)) cend = kmax >> 3;
  if (!active) cend = 0;
  float csn[CMAX + 1], csl[CMAX + 1];  // running sums of the aggregates of chunks tid + 1 .. tid + j
  csn[0] = csl[0] = 0.f;
#pragma unroll 1
  for (int cc = 1; cc <= cend; ++cc) {
    csn[cc] = csn[cc - 1] + agn[c.tid + cc];
    csl[cc] = csl[cc - 1] + agl[c.tid + cc];
  }
  float zs0 = 0.f, zs1 = 0.f;
  // the positions of chunk tid + cc, walked downwards.  ALL: every (position, output) pair is a candidate; CLIP: the
  // chunk straddles the horizon; ARITH: some l <= 0 before the chunk, every candidate of it has an arithmetic prefix
  auto chunk = [&](int cc, auto all, auto clip, auto arith, float (&zc0)[8], float (&zc1)[8], float (&rc)[8]) {
    float l[8], r[8], Ln[8], Ll[8], ut[15];
    load8(lpl, c.tid + cc, l);
    load8(rpl, c.tid + cc, r);
    {
      float sn = csn[cc - 1], sl = csl[cc - 1];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        sn += fminf(l[k], 0.f);
        Ln[k] = sn;
        if constexpr (!decltype(arith)::value) {
          sl += l[k] <= 0.f ? f_inf() : fast_lg2(1.f + l[k]);
          Ll[k] = sl;
        }
      }
    }
    // w0 / (sw cnt) for cnt = 8 cc + (k - i) + 1, k - i in [-7, 7]
#pragma unroll
    for (int j = 0; j < 15; ++j) ut[j] = w0p * fast_rcp((float)(8 * cc + j - 6));
#pragma unroll
    for (int k = 7; k >= 0; --k) {
      float rs = 0.f, q0 = 0.f, q1 = 0.f;
      if (!decltype(clip)::value || 8 * cc + k <= kmax) {
        const float rk = r[k];
        const bool rpos = rk > 0.f;
        const float rnp = w1p * fminf(rk, 0.f);
        if constexpr (decltype(arith)::value) {
          float sdz = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (decltype(all)::value || (8 * cc + k - i >= UA && 8 * cc + k - i <= UB)) {
              float d0, dz;
              cand_arith(i, hneg[i] + Ln[k], ut[k - i + 7], rnp, d0, dz);
              Z0[i] += d0;
              q0 += d0;
              sdz += dz;
            }
          rs = rpos ? 0.f : sdz;
        } else {
          const float rlp = w1p * fast_lg2(1.f + rk);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (decltype(all)::value || (8 * cc + k - i >= UA && 8 * cc + k - i <= UB)) {
              float d0, d1, kr;
              cand(i, hneg[i] + Ln[k], hlog[i] + Ll[k], ut[k - i + 7], rnp, rlp, rpos, d0, d1, kr);
              Z0[i] += d0;
              Z1[i] += d1;
              q0 += d0;
              q1 += d1;
              rs += kr;
            }
        }
      }
      zs0 += q0;
      zs1 += q1;
      zc0[k] = zs0;
      zc1[k] = zs1;
      rc[k] = rs;
    }
  };
  auto step = [&](int cc, auto all) {
    float zc0[8], zc1[8], rc[8];
    if (cc <= cend) {
      const bool ar = csl[cc - 1] == f_inf();
      if (8 * cc + 7 <= kmax) {
        if (ar) chunk(cc, all, BoolC<false>{}, BoolC<true>{}, zc0, zc1, rc);
        else chunk(cc, all, BoolC<false>{}, BoolC<false>{}, zc0, zc1, rc);
      } else {
        chunk(cc, all, BoolC<true>{}, BoolC<false>{}, zc0, zc1, rc);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) zc0[k] = zc1[k] = rc[k] = 0.f;
    }
    if (cc == CMAX) {  // the first step initialises every slot: its writer's numbers, or zeros where there is no writer
      if (c.tid + cc < NT) {
        store8(g0p, c.tid + cc, zc0);
        store8(g1p, c.tid + cc, zc1);
        store8(grp, c.tid + cc, rc);
      }
      if (c.tid < CMAX) {
        plane_zero8(g0p, c.tid);
        plane_zero8(g1p, c.tid);
        plane_zero8(grp, c.tid);
      }
    } else if (c.tid + cc < NT && cc <= cend) {  // one writer per chunk and step: a plain read-modify-write
      float a[8], b[8], d[8];
      load8(g0p, c.tid + cc, a);
      load8(g1p, c.tid + cc, b);
      load8(grp, c.tid + cc, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a[k] += zc0[k];
        b[k] += zc1[k];
        d[k] += rc[k];
      }
      store8(g0p, c.tid + cc, a);
      store8(g1p, c.tid + cc, b);
      store8(grp, c.tid + cc, d);
    }
    simt::sync();
  };
  constexpr int CF0 = (UA + 7 + 7) / 8 > 1 ? (UA + 7 + 7) / 8 : 1;  // first chunk whose every pair is a candidate
  constexpr int CF1 = UB >= 15 ? (UB - 7) / 8 : 0;                   // last such chunk (0: none)
  if constexpr (CF1 >= CF0) {
#pragma unroll
    for (int cc = CMAX; cc > CF1; --cc) step(cc, BoolC<false>{});
#pragma unroll 1
    for (int cc = CF1; cc >= CF0; --cc) step(cc, BoolC<true>{});
#pragma unroll
    for (int cc = CF0 - 1; cc >= 1; --cc) step(cc, BoolC<false>{});
  } else {
#pragma unroll
    for (int cc = CMAX; cc >= 1; --cc) step(cc, BoolC<false>{});
  }
  if (live) {  // what the threads before this one handed over
    float a[8], b[8], d[8];
    load8(g0p, c.tid, a);
    load8(g1p, c.tid, b);
    load8(grp, c.tid, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      z0_own[k] += a[k];
      z1_own[k] += b[k];
      r_own[k] += d[k];
    }
  }
  if (!active) return;
  // own chunk: triangular, in registers
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float t0[8], t1[8];
    float pn = 0.f, pl = kg[i] != 0.f ? 0.f : f_inf();
#pragma unroll
    for (int q = i; q < 8; ++q) {
      pn += lnq[q];
      pl += llq[q];
      t0[q] = t1[q] = 0.f;
      if (q - i >= UA && q - i <= UB && q <= kmax) {
        float kr;
        cand(i, pn, pl, w0p * (1.f / (float)(q - i + 1)), w1p * fminf(ro[q], 0.f), w1p * fast_lg2(1.f + ro[q]), ro[q] > 0.f, t0[q], t1[q], kr);
        r_own[q] += kr;
      }
    }
    float run0 = Z0[i], run1 = Z1[i];
#pragma unroll
    for (int q = 7; q >= i; --q) {
      run0 += t0[q];
      run1 += t1[q];
      z0_own[q] += run0;
      z1_own[q] += run1;
    }
  }
}
// owner: the factors in l_k and r_q (Adj<SEM_AGM, MODE_AND>::in_factor), routed through the operands' negations
template <bool NL, bool NR>
PF_DEV void agm_until_adj_finish(const float* lpl, const float* rpl, int tid, float w0, float w1, const float (&z0)[8],
                                 const float (&z1)[8], const float (&rr)[8], float (&gsl)[8], float (&gsr)[8]) {
  const float w1p = w1 * fast_rcp(w0 + w1);
  float l[8], r[8];
  load8(lpl, tid, l);
  load8(rpl, tid, r);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float dl = l[i] > 0.f ? z1[i] * fast_rcp(1.f + l[i]) : (l[i] < 0.f ? z0[i] : 0.5f * z0[i]);
    const float dr = w1p * rr[i] * (r[i] > 0.f ? fast_rcp(1.f + r[i]) : (r[i] < 0.f ? 1.f : 0.5f));
    gsl[i] += NL ? -dl : dl;
    gsr[i] += NR ? -dr : dr;
  }
}

// ---- dgmsr until, backward, FAST PATH (p = 1, unit weights_pair, uniform weights_left / weights_right): the adjoint of
// UntilOp::phase2_dgmsr on the lockstep walk, with the chain of Adj<SEM_DGMSR> partials written out per case
// (eps: rp.eps_p, s = sqrt(eps); a0 / a1: the prefix accumulator, sum of l^2 over l <= 0 / sum of 2 lg2 l or +inf):
//   Or (U = -And(-z) over n candidates):  U >= 0:  g dU/dz = kA z [z >= 0],  kA = g / (n (U + s));
//                                          U <  0:  g dU/dz = kG / (-z),      kG = g (s - U) (1 - (eps / (s - U)^2)^n) / n;
//   flagged prefix (some l <= 0):  sy = sqrt(eps + a0 / cnt), y = s - sy <= 0, sz = sqrt(eps + (y^2 + min(r, 0)^2) / 2), z = s - sz:
//        half = dz / (2 sz);   r-route  half |r| [r <= 0];   l-route  Z0 += half (sy - s) / (cnt sy),  applied as |l_k| Z0 [l_k <= 0];
//   all l > 0, r <= 0:  z depends on the position only and y carries no weight: r-route  dz |r| / (2 sz_r), no l-route;
//   all l > 0, r >  0:  ey = 2^(lae / (2 cnt)) (lae = logaddexp2(cnt lg2 eps, a1)), y = ey - s, q = y r, m2 = eps^2 + q^2,
//        ez = m2^(1/4), z = ez - s:   c = dz ez q / (2 m2);   r-route  c y;   l-route  Z1 += c r ey (1 - 2^(cnt lg2 eps - lae)) / cnt,
//        applied as Z1 / l_k [l_k > 0].
// A chunk behind some l <= 0 of the shared tail holds flagged prefixes only (five transcendentals per candidate).
template <int NT>
PF_DEV void dgmsr_until_adj_stage(const TC& c, const float* lpl, float* ag0, float* ag1) {
  float l[8];
  load8(lpl, c.tid, l);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (c.p0 + i < c.nlive) {
      const float ln = fminf(l[i], 0.f);
      s0 = fmaf(ln, ln, s0);
      s1 += l[i] <= 0.f ? f_inf() : 2.f * fast_lg2(l[i]);
    }
  ag0[c.tid] = s0;
  ag1[c.tid] = s1;
}
// kA_i / kG_i: the output-level coefficients of the Or (one of them is zero)
template <int UA, int UB>
PF_DEV void dgmsr_until_adj_coeff(const TC& c, const RP& rp, const float (&g)[8], const Vec<1> (&U)[8], float (&kA)[8], float (&kG)[8]) {
  const int kmax = c.nlive - 1 - c.p0;
  const float se = rp.sqrt_eps, lge = rp.ln_eps * kLog2e;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int last = kmax - i;
    if (last > UB) last = UB;
    const int n = last - UA + 1;
    const bool has = g[i] != 0.f && n > 0;
    const float nf = (float)(n > 0 ? n : 1);
    const float u = U[i].x[0];
    const float gn = g[i] * fast_rcp(nf);
    const float rm0 = se - u;  // sqrt(m0) of the geometric Or
    const float ratio = fast_ex2(nf * (lge - 2.f * fast_lg2(fmaxf(rm0, 1e-30f))));
    kA[i] = (has && u >= 0.f) ? gn * fast_rcp(u + se) : 0.f;
    kG[i] = (has && !(u >= 0.f)) ? gn * rm0 * (1.f - ratio) : 0.f;
  }
}
template <int UA, int UB, int NT>
PF_DEV void dgmsr_until_adj_walk(const TC& c, bool live, const RP& rp, const float* lpl, const float* rpl, const float* ag0,
                                 const float* ag1, float* g0p, float* g1p, float* grp, const float (&kA)[8], const float (&kG)[8],
                                 float (&z0_own)[8], float (&z1_own)[8], float (&r_own)[8]) {
  static_assert(UA >= 0 && UB >= UA, "bounded until");
  constexpr int CMAX = (UB + 7) / 8;
  const int kmax = c.interior ? 0x3fffffff : c.nlive - 1 - c.p0;
  const float eps = rp.eps_p, se = rp.sqrt_eps, lge = rp.ln_eps * kLog2e;
  bool active = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) active = active || kA[i] != 0.f || kG[i] != 0.f;
  active = active && live;
  float lo[8], ro[8], l2q[8], llq[8], h0[8], h1[8], Z0[8], Z1[8];
  load8(lpl, c.tid, lo);
  load8(rpl, c.tid, ro);
  {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      const bool ok = c.p0 + i < c.nlive;
      const float ln = fminf(lo[i], 0.f);
      l2q[i] = ok ? ln * ln : 0.f;
      llq[i] = ok ? (lo[i] <= 0.f ? f_inf() : 2.f * fast_lg2(lo[i])) : 0.f;
      s0 += l2q[i];
      s1 += llq[i];
      h0[i] = s0;
      h1[i] = s1;
      Z0[i] = Z1[i] = 0.f;
    }
  }
  // flagged candidate of output i: a0 = sum of l^2 over the l <= 0 of the prefix, invk = 1 / cnt, r2n = min(r, 0)^2
  //   -> d0: contribution to Z0; kr: to the r-route (the owner multiplies by |r|; nothing when r > 0)
  auto cand_flag = [&](int i, float a0, float invk, float r2n, bool rpos, float& d0, float& kr) {
    const float sy = fast_sqrt(fmaf(a0, invk, eps));
    const float y = se - sy;
    const float sz = fast_sqrt(fmaf(0.5f, fmaf(y, y, r2n), eps));
    const float nz = sz - se;  // -z >= 0
    const float dz = nz > 0.f ? kG[i] * fast_rcp(nz) : 0.f;  // (z == 0 joins an arithmetic Or with weight z = 0)
    const float half = 0.5f * dz * fast_rcp(sz);
    kr = rpos ? 0.f : half;
    d0 = half * (sy - se) * invk * fast_rcp(sy);
  };
  // general candidate; szr / crz: sqrt(eps + r^2 / 2) and 1 / (2 szr (szr - s)) of this position (r <= 0)
  auto cand = [&](int i, float a0, float a1, float invk, float kf, float r, float r2n, float szr, float crz, float& d0, float& d1, float& kr) {
    const bool flag = a1 == f_inf();
    float d0f, krf;
    cand_flag(i, a0, invk, r2n, r > 0.f, d0f, krf);
    // all l > 0
    const float u = kf * lge;
    const float lae = fmaxf(u, a1) + fast_lg2(1.f + fast_ex2(-fabsf(u - a1)));
    const float ey = fast_ex2(0.5f * invk * lae);
    const float y = ey - se;
    const float q = y * r;
    const float m2 = fmaf(q, q, eps * eps);
    const float ez = fast_sqrt(fast_sqrt(m2));
    const float z = ez - se;
    const float cg = 0.5f * (kA[i] * z) * ez * q * fast_rcp(m2);  // (z > 0 belongs to an arithmetic Or: g dU/dz = kA z)
    const float ry = 1.f - fast_ex2(u - lae);
    const float d1g = cg * r * ey * ry * invk;
    // all l > 0, r <= 0: z = s - szr < 0 (or 0 when r == 0), only the r-route
    const float krn = (szr > se) ? kG[i] * crz : 0.f;
    const bool rpos = r > 0.f;
    d0 = flag ? d0f : 0.f;
    d1 = (!flag && rpos) ? d1g : 0.f;
    kr = flag ? krf : (rpos ? cg * y : krn);
  };
  int cend = CMAX;
  if (cend > NT - 1 - c.tid) cend = NT - 1 - c.tid;
  if (cend > (kmax >> 3)) cend = kmax >> 3;
  if (!active) cend = 0;
  float cs0[CMAX + 1], cs1[CMAX + 1];  // running sums of the aggregates of chunks tid + 1 .. tid + j
  cs0[0] = cs1[0] = 0.f;
#pragma unroll 1
  for (int cc = 1; cc <= cend; ++cc) {
    cs0[cc] = cs0[cc - 1] + ag0[c.tid + cc];
    cs1[cc] = cs1[cc - 1] + ag1[c.tid + cc];
  }
  float zs0 = 0.f, zs1 = 0.f;
  auto chunk = [&](int cc, auto all, auto clip, auto flagged, float (&zc0)[8], float (&zc1)[8], float (&rc)[8]) {
    float l[8], r[8], L0[8], L1[8], it[15];
    load8(lpl, c.tid + cc, l);
    load8(rpl, c.tid + cc, r);
    {
      float s0 = cs0[cc - 1], s1 = cs1[cc - 1];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float ln = fminf(l[k], 0.f);
        s0 = fmaf(ln, ln, s0);
        L0[k] = s0;
        if constexpr (!decltype(flagged)::value) {
          s1 += l[k] <= 0.f ? f_inf() : 2.f * fast_lg2(l[k]);
          L1[k] = s1;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 15; ++j) it[j] = fast_rcp((float)(8 * cc + j - 6));  // 1 / cnt, cnt = 8 cc + (k - i) + 1
#pragma unroll
    for (int k = 7; k >= 0; --k) {
      float rs = 0.f, q0 = 0.f, q1 = 0.f;
      if (!decltype(clip)::value || 8 * cc + k <= kmax) {
        const float rk = r[k];
        const float rn = fminf(rk, 0.f);
        const float r2n = rn * rn;
        if constexpr (decltype(flagged)::value) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (decltype(all)::value || (8 * cc + k - i >= UA && 8 * cc + k - i <= UB)) {
              float d0, kr;
              cand_flag(i, h0[i] + L0[k], it[k - i + 7], r2n, rk > 0.f, d0, kr);
              Z0[i] += d0;
              q0 += d0;
              rs += kr;
            }
        } else {
          const float szr = fast_sqrt(fmaf(0.5f, r2n, eps));
          const float crz = 0.5f * fast_rcp(szr * fmaxf(szr - se, 1e-30f));
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (decltype(all)::value || (8 * cc + k - i >= UA && 8 * cc + k - i <= UB)) {
              float d0, d1, kr;
              cand(i, h0[i] + L0[k], h1[i] + L1[k], it[k - i + 7], (float)(8 * cc + k - i + 1), rk, r2n, szr, crz, d0, d1, kr);
              Z0[i] += d0;
              Z1[i] += d1;
              q0 += d0;
              q1 += d1;
              rs += kr;
            }
        }
      }
      zs0 += q0;
      zs1 += q1;
      zc0[k] = zs0;
      zc1[k] = zs1;
      rc[k] = rs;
    }
  };
  auto step = [&](int cc, auto all) {
    float zc0[8], zc1[8], rc[8];
    if (cc <= cend) {
      const bool fl = cs1[cc - 1] == f_inf();
      if (8 * cc + 7 <= kmax) {
        if (fl) chunk(cc, all, BoolC<false>{}, BoolC<true>{}, zc0, zc1, rc);
        else chunk(cc, all, BoolC<false>{}, BoolC<false>{}, zc0, zc1, rc);
      } else {
        chunk(cc, all, BoolC<true>{}, BoolC<false>{}, zc0, zc1, rc);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) zc0[k] = zc1[k] = rc[k] = 0.f;
    }
    if (cc == CMAX) {  // the first step initialises every slot: its writer's numbers, or zeros where there is no writer
      if (c.tid + cc < NT) {
        store8(g0p, c.tid + cc, zc0);
        store8(g1p, c.tid + cc, zc1);
        store8(grp, c.tid + cc, rc);
      }
      if (c.tid < CMAX) {
        plane_zero8(g0p, c.tid);
        plane_zero8(g1p, c.tid);
        plane_zero8(grp, c.tid);
      }
    } else if (c.tid + cc < NT && cc <= cend) {  // one writer per chunk and step: a plain read-modify-write
      float a[8], b[8], d[8];
      load8(g0p, c.tid + cc, a);
      load8(g1p, c.tid + cc, b);
      load8(grp, c.tid + cc, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a[k] += zc0[k];
        b[k] += zc1[k];
        d[k] += rc[k];
      }
      store8(g0p, c.tid + cc, a);
      store8(g1p, c.tid + cc, b);
      store8(grp, c.tid + cc, d);
    }
    simt::sync();
  };
  constexpr int CF0 = (UA + 7 + 7) / 8 > 1 ? (UA + 7 + 7) / 8 : 1;
  constexpr int CF1 = UB >= 15 ? (UB - 7) / 8 : 0;
  if constexpr (CF1 >= CF0) {
#pragma unroll
    for (int cc = CMAX; cc > CF1; --cc) step(cc, BoolC<false>{});
#pragma unroll 1
    for (int cc = CF1; cc >= CF0; --cc) step(cc, BoolC<true>{});
#pragma unroll
    for (int cc = CF0 - 1; cc >= 1; --cc) step(cc, BoolC<false>{});
  } else {
#pragma unroll
    for (int cc = CMAX; cc >= 1; --cc) step(cc, BoolC<false>{});
  }
  if (live) {  // what the threads before this one handed over
    float a[8], b[8], d[8];
    load8(g0p, c.tid, a);
    load8(g1p, c.tid, b);
    load8(grp, c.tid, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      z0_own[k] += a[k];
      z1_own[k] += b[k];
      r_own[k] += d[k];
    }
  }
  if (!active) return;
  // own chunk: triangular, in registers
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float t0[8], t1[8];
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int q = i; q < 8; ++q) {
      a0 += l2q[q];
      a1 += llq[q];
      t0[q] = t1[q] = 0.f;
      if (q - i >= UA && q - i <= UB && q <= kmax) {
        const float rn = fminf(ro[q], 0.f);
        const float r2n = rn * rn;
        const float szr = fast_sqrt(fmaf(0.5f, r2n, eps));
        const float crz = 0.5f * fast_rcp(szr * fmaxf(szr - se, 1e-30f));
        float kr;
        cand(i, a0, a1, 1.f / (float)(q - i + 1), (float)(q - i + 1), ro[q], r2n, szr, crz, t0[q], t1[q], kr);
        r_own[q] += kr;
      }
    }
    float run0 = Z0[i], run1 = Z1[i];
#pragma unroll
    for (int q = 7; q >= i; --q) {
      run0 += t0[q];
      run1 += t1[q];
      z0_own[q] += run0;
      z1_own[q] += run1;
    }
  }
}
// owner: the factors in l_k and r_q (Adj<SEM_DGMSR, MODE_AND>::in_factor with p = 1), routed through the operands' negations
template <bool NL, bool NR>
PF_DEV void dgmsr_until_adj_finish(const float* lpl, const float* rpl, int tid, const float (&z0)[8], const float (&z1)[8],
                                   const float (&rr)[8], float (&gsl)[8], float (&gsr)[8]) {
  float l[8], r[8];
  load8(lpl, tid, l);
  load8(rpl, tid, r);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float dl = l[i] <= 0.f ? z0[i] * -l[i] : z1[i] * fast_rcp(l[i]);
    const float dr = r[i] <= 0.f ? rr[i] * -r[i] : rr[i];
    gsl[i] += NL ? -dl : dl;
    gsr[i] += NR ? -dr : dr;
  }
}

// ---- agm / dgmsr until, backward (agm.py:90-130, dgmsr.py:88-118; optional weights).
//   y_i = AND_{k<=i}(l_k; wl),  z_i = AND((y_i, r_i); wp),  U = OR_{i>=UA}(z_i; wr)
// Sweep 1 recomputes y_i, z_i (kept in a per-thread scratch array) and U; sweep 2 walks i downwards:
//   dU/dr_i = dOR/dz_i * dz_i/dr_i,   dU/dl_k = sum_{i >= max(k, UA)} dOR/dz_i * dz_i/dy_i * dy_i/dl_k,
// where dy_i/dl_k = in_factor(out_coeff(y_i), l_k) factorises, so the sum over i is a running
// suffix sum of the two-word coefficient -- no subtraction anywhere.
template <int SEM, int UA, int UB, int NT>
PF_DEV void gm_until_scatter(const TC& c, const RP& rp, float wp0, float wp1, const float* wl, const float* wr, const float (&g)[8],
                             const float* lpl, const float* rpl, float* gl, float* gr) {
  using RA = Red<SEM, MODE_AND>;
  using RO = Red<SEM, MODE_OR>;
  using AA = Adj<SEM, MODE_AND>;
  using AO = Adj<SEM, MODE_OR>;
  float ybuf[UB + 1], zbuf[UB + 1];
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    float gi = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j == i) gi = g[j];
    const int p = c.p0 + i;
    int last = c.nlive - 1 - p;
    if (last > UB) last = UB;
    if (gi == 0.f || last < UA) continue;
    // sweep 1
    Vec<2> pre = RA::ident(), orr = RO::ident();
    float swl = 0.f, swr = 0.f;
    for (int kk = 0; kk <= last; ++kk) {
      const int q = pidx(p + kk);
      const float wlk = wl ? wl[kk] : 1.f;
      pre = RA::comb(pre, RA::lift(Vec<1>{{lpl[q]}}, rp, wlk));
      swl += wlk;
      if (kk >= UA) {
        const Vec<1> y = RA::fin(pre, swl, rp);
        const Vec<1> z = RA::fin(RA::comb(RA::lift(y, rp, wp0), RA::lift(Vec<1>{{rpl[q]}}, rp, wp1)), wp0 + wp1, rp);
        const float wrj = wr ? wr[kk - UA] : 1.f;
        orr = RO::comb(orr, RO::lift(z, rp, wrj));
        swr += wrj;
        ybuf[kk] = y.x[0];
        zbuf[kk] = z.x[0];
      }
    }
    if (!(swr > 0.f)) continue;
    const float U = RO::fin(orr, swr, rp).x[0];
    bool lost;
    const Vec<2> kor = AO::out_coeff(gi, U, swr, rp, lost);
    // sweep 2: suffix sum of a_i * out_coeff(y_i)
    Vec<2> suf{{0.f, 0.f}};
    for (int kk = last; kk >= 0; --kk) {
      const int q = pidx(p + kk);
      if (kk >= UA) {
        const float wrj = wr ? wr[kk - UA] : 1.f;
        const float dz = AO::in_factor(kor, zbuf[kk], wrj, rp);  // dU/dz_i * g
        const Vec<2> kp = AA::out_coeff(dz, zbuf[kk], wp0 + wp1, rp, lost);
        simt::atomic_add(&gr[q], AA::in_factor(kp, rpl[q], wp1, rp));
        const float a = AA::in_factor(kp, ybuf[kk], wp0, rp);  // dU/dy_i * g
        float swl_i = 0.f;  // sum of the prefix weights up to kk
        if (wl) {
          for (int k2 = 0; k2 <= kk; ++k2) swl_i += wl[k2];
        } else {
          swl_i = (float)(kk + 1);
        }
        const Vec<2> ky = AA::out_coeff(a, ybuf[kk], swl_i, rp, lost);
        suf.x[0] += ky.x[0];
        suf.x[1] += ky.x[1];
      }
      simt::atomic_add(&gl[q], AA::in_factor(suf, lpl[q], wl ? wl[kk] : 1.f, rp));
    }
  }
}

// g_src (+)= the thread's chunk of a scatter plane
template <bool NEG>
PF_DEV void accumulate8(const float* plane, int tid, float (&gs)[8]) {
  float t[8];
  load8(plane, tid, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) gs[i] += NEG ? -t[i] : t[i];
}

// ====================================================================== cumulative backward
// Cumulative values are (pos, neg) pairs; And / Always are per-channel minima, Or maxima,
// Eventually a SUM, Until a sum of minima (cumulative.py:35-103).  The adjoint of a sum window is
// a backward-looking window sum like everywhere else; minima / maxima route their cotangent to ONE
// operand, the first extremal one, like torch.min / max(dim) in the reference
// (torch/semantics.py:253-275): the backward kernel evaluates those windows on (value, position)
// pairs and scatters the cotangent through shared memory.
template <int MODE>
struct ArgRed {
  using A = Vec<2>;  // (value, bits of the tile position)
  static PYSTL_HD A ident() { return A{{MODE == MODE_AND ? f_inf() : -f_inf(), 0.f}}; }
  static PYSTL_HD A comb(A a, A b) {  // a is the EARLIER operand: it keeps ties
    const bool tb = MODE == MODE_AND ? (b.x[0] < a.x[0]) : (b.x[0] > a.x[0]);
    return A{{tb ? b.x[0] : a.x[0], tb ? b.x[1] : a.x[1]}};
  }
};
template <int MODE>
PF_DEV void arg_lift8(const TC& c, const Vec<2> (&v)[8], int w, Vec<2> (&x)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (c.p0 + i < c.nlive) ? Vec<2>{{v[i].x[w], i2f(c.p0 + i)}} : ArgRed<MODE>::ident();
}
// cotangent of the thread's outputs -> the positions their extremum came from
PF_DEV void arg_scatter8(float* plane, const Vec<2> (&acc)[8], const Vec<2> (&g)[8], int w) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (g[i].x[w] != 0.f) simt::atomic_add(&plane[pidx(f2i(acc[i].x[1]))], g[i].x[w]);
}
// g_src (+)= the thread's chunk of the scatter planes; NEG: the operand was Not(src) = (-neg, -pos)
template <bool NEG>
PF_DEV void cum_accumulate8(const float* plane_pos, const float* plane_neg, int tid, Vec<2> (&gs)[8]) {
  float a[8], b[8];
  load8(plane_pos, tid, a);
  load8(plane_neg, tid, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    gs[i].x[0] += NEG ? -b[i] : a[i];
    gs[i].x[1] += NEG ? -a[i] : b[i];
  }
}
// the same from registers (the gathered r-routes of the Until adjoint): a = pos channel, b = neg channel
template <bool NEG>
PF_DEV void cum_add_ab8(const float (&a)[8], const float (&b)[8], Vec<2> (&gs)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    gs[i].x[0] += NEG ? -b[i] : a[i];
    gs[i].x[1] += NEG ? -a[i] : b[i];
  }
}
template <bool NEG>
PF_DEV void cum_add8(const Vec<2> (&w)[8], Vec<2> (&gs)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    gs[i].x[0] += NEG ? -w[i].x[1] : w[i].x[0];
    gs[i].x[1] += NEG ? -w[i].x[0] : w[i].x[1];
  }
}
// n-ary min / max of one sample: the cotangent goes to the first extremal operand, per channel
template <int MODE, int N>
PF_DEV void cum_nary_bwd(const Vec<2> (&a)[N], Vec<2> g, Vec<2> (&ga)[N]) {
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    int best = 0;
    float bv = a[0].x[w];
#pragma unroll
    for (int j = 1; j < N; ++j) {
      const bool t = MODE == MODE_AND ? (a[j].x[w] < bv) : (a[j].x[w] > bv);
      bv = t ? a[j].x[w] : bv;
      best = t ? j : best;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) ga[j].x[w] = (j == best) ? g.x[w] : 0.f;
  }
}
// root cotangent of a cumulative formula: grad_out[b, t, 2] or the constant for both channels
template <bool NEG>
PF_DEV void root_cotangent2(const TC& c, long long T, const float* __restrict__ grad_out, float scale, Vec<2> (&g)[8]) {
  const float* src = grad_out ? grad_out + (c.b * T + c.origin + c.p0) * 2 : nullptr;
  int hi = c.p_hi;
  if (hi > c.last_ok + 1) hi = c.last_ok + 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a = 0.f, b = 0.f;
    if (c.p0 + i < hi) {
      a = src ? src[2 * i] : scale;
      b = src ? src[2 * i + 1] : scale;
    }
    g[i] = NEG ? Vec<2>{{-b, -a}} : Vec<2>{{a, b}};
  }
}
// predicate (pos, neg) = (max(0, v), min(0, v)): d/dv as torch.maximum / minimum give it
PF_DEV float cum_pred_adj(Vec<2> v, Vec<2> g) {
  const float x = v.x[0] + v.x[1];
  return g.x[0] * dmax0(x) + g.x[1] * dmin0(x);
}
// Until: U = sum_{i in [UA, last]} min(r_i, min_{k<=i} l_k) per channel (cumulative.py:96-103).
// Each candidate routes the cotangent to r_i or to the first prefix minimum of l (ties between
// the two split, torch.minimum).  lpl / rpl: staged operands; gl / gr: zeroed scatter planes.
// ---- adjoint of the cumulative Until, U_t = sum_{i in [UA, UB]} min(r_{t+i}, P_t(t+i)), P_t(q) = min l[t..q]
// (cumulative.py:96-103; torch.minimum routes the cotangent of a candidate to r_q when r_q < P, to the
// first prefix minimum of l when P < r_q, and splits it on a tie).
//
// r-routes as a GATHER: grad_r[q] = sum_{t in [q - UB, q - UA]} g_t [r_q < P_t(q)] lands in the
// registers of the thread that owns q -- no atomics.  For an output t before the thread's chunk,
// P_t(q) = min(Mb_t, hf_q): Mb_t = min l[t .. p0 - 1] is one running minimum of the backward walk,
// hf_q = min l[p0 .. q] a prefix minimum of the own chunk.  gpl: the cotangent plane (store8 layout,
// zero where no output exists).
PF_DEV float cum_route_r(float r, float P, float g) { return r < P ? g : (r == P ? 0.5f * g : 0.f); }
template <int UA, int UB, int NT>
PF_DEV void cum_until_gather_r(const TC& c, const Vec<2> (&g)[8], int w, const float* lpl, const float* rpl, const float* gpl,
                               float (&racc)[8]) {
  const int f0 = c.p0 >> 2;
  const F4* L4 = reinterpret_cast<const F4*>(lpl);
  const F4* R4 = reinterpret_cast<const F4*>(rpl);
  const F4* G4 = reinterpret_cast<const F4*>(gpl);
  const F4 la = L4[pf4(f0)], lb = L4[pf4(f0 + 1)], ra = R4[pf4(f0)], rb = R4[pf4(f0 + 1)];
  const float lo[8] = {la.v[0], la.v[1], la.v[2], la.v[3], lb.v[0], lb.v[1], lb.v[2], lb.v[3]};
  const float ro[8] = {ra.v[0], ra.v[1], ra.v[2], ra.v[3], rb.v[0], rb.v[1], rb.v[2], rb.v[3]};
  float hf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    racc[j] = 0.f;
    hf[j] = j ? fminf(hf[j - 1], lo[j]) : lo[0];
  }
  // outputs of the own chunk
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float pm = f_inf();
#pragma unroll
    for (int q = i; q < 8; ++q) {
      pm = fminf(pm, lo[q]);
      if (q - i >= UA && q - i <= UB) racc[q] += cum_route_r(ro[q], pm, g[i].x[w]);
    }
  }
  // outputs before it: t = p0 - m, m = 1 .. UB; position q = p0 + j is candidate m + j of output t
  float Mb = f_inf();
  constexpr int G0 = UA > 0 ? (UA - 1 + 3) / 4 : 0;            // first group whose every (m, j) has m + j >= UA ... m >= UA
  constexpr int G1 = UB >= 11 ? (UB - 11) / 4 + 1 : 0;         // groups below G1 have m + 7 <= UB
  constexpr int GB0 = G0 < G1 ? G0 : G1, GB1 = G1;             // body [GB0, GB1)
  constexpr int GN = (UB + 3) / 4;                             // groups holding m <= UB
  auto group = [&](int gg, auto all) {
    const int fb = f0 - 1 - gg;
    const bool ok = fb >= 0;  // (outputs before the tile do not exist for this CTA)
    const F4 ql = L4[pf4(ok ? fb : 0)];
    F4 qg = G4[pf4(ok ? fb : 0)];
#pragma unroll
    for (int k = 3; k >= 0; --k) {
      const int m = 4 * gg + (4 - k);
      const float gt = ok ? qg.v[k] : 0.f;
      Mb = fminf(Mb, ql.v[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (decltype(all)::value || (m + j >= UA && m + j <= UB)) racc[j] += cum_route_r(ro[j], fminf(Mb, hf[j]), gt);
    }
  };
#pragma unroll
  for (int gg = 0; gg < GB0; ++gg) group(gg, BoolC<false>{});
#pragma unroll 1
  for (int gg = GB0; gg < GB1; ++gg) group(gg, BoolC<true>{});
#pragma unroll
  for (int gg = GB1 > GB0 ? GB1 : GB0; gg < GN; ++gg) group(gg, BoolC<false>{});
}
// l-routes as a SCATTER with the runs merged, the adjoint of UntilOp::phase2_cum: the 8 outputs of a
// thread walk their candidates together.  Past the own chunk the prefix minimum of output i at
// position q is min(h_i, M_q) (h_i is the earlier one: it wins ties); the cotangent routed to a
// prefix minimum stays in a register while that minimum stands: one shared-memory atomic per record
// low of M and one per output for its own-chunk minimum, instead of one per candidate.
template <int UA, int UB, int NT>
PF_DEV void cum_until_scatter_l_fast(const TC& c, const Vec<2> (&g)[8], int w, const float* lpl, const float* rpl, float* gl) {
  constexpr int LEXT = 8 * NT, FMAX = LEXT / 4 - 1;
  const int p0 = c.p0, f0 = p0 >> 2;
  const int kmax = c.interior ? 0x3fffffff : c.nlive - 1 - p0;  // windows are clipped at the horizon on the last tile
  const F4* L4 = reinterpret_cast<const F4*>(lpl);
  const F4* R4 = reinterpret_cast<const F4*>(rpl);
  float gi[8], h[8], hacc[8];
  int harg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) gi[i] = g[i].x[w];
  {  // own chunk: triangular
    const F4 la = L4[pf4(f0)], lb = L4[pf4(f0 + 1)], ra = R4[pf4(f0)], rb = R4[pf4(f0 + 1)];
    const float lo[8] = {la.v[0], la.v[1], la.v[2], la.v[3], lb.v[0], lb.v[1], lb.v[2], lb.v[3]};
    const float ro[8] = {ra.v[0], ra.v[1], ra.v[2], ra.v[3], rb.v[0], rb.v[1], rb.v[2], rb.v[3]};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float pm = f_inf(), run = 0.f;
      int arg = p0 + i;
#pragma unroll
      for (int q = i; q < 8; ++q) {
        if (q > kmax) break;
        if (lo[q] < pm) {  // a new prefix minimum (first index wins ties): settle the old one
          if (run != 0.f) simt::atomic_add(&gl[pidx(arg)], run);
          run = 0.f;
          pm = lo[q];
          arg = p0 + q;
        }
        if (q - i >= UA && q - i <= UB) run += gi[i] - cum_route_r(ro[q], pm, gi[i]);
      }
      h[i] = pm;
      harg[i] = arg;
      hacc[i] = run;
    }
  }
  float M = f_inf(), accM = 0.f;
  int argM = p0 + 8;
  constexpr int KS0 = UA + 7 > 8 ? UA + 7 : 8;
  constexpr int KS = (KS0 + 3) & ~3;
  constexpr int KE = ((UB + 1) & ~3) > KS ? ((UB + 1) & ~3) : KS;  // body [KS, KE): every output takes every position
  auto group = [&](int gq, auto all) {
    const int fc = (f0 + gq) < FMAX ? (f0 + gq) : FMAX;
    const F4 ql = L4[pf4(fc)], qr = R4[pf4(fc)];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int kq = 4 * gq + k;
      if (kq > kmax) break;
      const float l = ql.v[k], r = qr.v[k];
      if (l < M) {
        if (accM != 0.f) simt::atomic_add(&gl[pidx(argM)], accM);
        accM = 0.f;
        M = l;
        argM = p0 + kq;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (decltype(all)::value || (kq - i >= UA && kq - i <= UB)) {
          const bool own = h[i] <= M;
          const float tl = gi[i] - cum_route_r(r, own ? h[i] : M, gi[i]);
          hacc[i] += own ? tl : 0.f;
          accM += own ? 0.f : tl;
        }
    }
  };
#pragma unroll
  for (int gq = 2; gq < KS / 4; ++gq) group(gq, BoolC<false>{});
#pragma unroll 1
  for (int gq = KS / 4; gq < KE / 4; ++gq) group(gq, BoolC<true>{});
#pragma unroll
  for (int gq = KE / 4; gq <= (UB + 7) / 4; ++gq) group(gq, BoolC<false>{});
  if (accM != 0.f) simt::atomic_add(&gl[pidx(argM)], accM);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (hacc[i] != 0.f) simt::atomic_add(&gl[pidx(harg[i])], hacc[i]);
}
// the l-routes of one channel: fast path when every window of the thread is complete (or on an
// interior tile, where the outputs whose windows leave it have a zero cotangent), else output by output
template <int UA, int UB, int NT>
PF_DEV void cum_until_scatter_l(const TC& c, const Vec<2> (&g)[8], int w, const float* lpl, const float* rpl, float* gl) {
  if constexpr (UA <= 64) {
    cum_until_scatter_l_fast<UA, UB, NT>(c, g, w, lpl, rpl, gl);
    return;
  }
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    float gi = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j == i) gi = g[j].x[w];
    if (gi == 0.f) continue;
    const int p = c.p0 + i;
    int last = c.nlive - 1 - p;
    if (last > UB) last = UB;
    float pm = f_inf(), run = 0.f;
    int arg = p;
    for (int kk = 0; kk <= last; ++kk) {
      const int q = pidx(p + kk);
      const float l = lpl[q];
      if (l < pm) {
        if (run != 0.f) simt::atomic_add(&gl[arg], run);
        run = 0.f;
        pm = l;
        arg = q;
      }
      if (kk >= UA) run += gi - cum_route_r(rpl[q], pm, gi);
    }
    if (run != 0.f) simt::atomic_add(&gl[arg], run);
  }
}

// ---- cumulative until, backward, LOCKSTEP walk (one channel; UA <= 64).  Same hand-off scheme as the smooth / agm walks.
// With P_i(q) = min(h_i, M_q) (h_i: minimum of l over the rest of the own chunk, M_q: running minimum of the shared tail) a
// candidate routes its cotangent g_i to r_q when r_q < P_i(q), else to the first argmin of l[t_i .. q].  ONE threshold per
// position decides all 8 outputs:  thr = min(r_q, M_q);
//     thr <  h_i:  the cotangent leaves the thread -- to r_q when r_q < M_q, else to the tail's argmin (then h_i > M_q): the
//                  SUM of those g_i is routed once per position; the tail's argmin is the last record low of l at or below q,
//                  and in a downward walk it is reached exactly when its amount is complete: flush at the record;
//     thr >= h_i:  h_i <= min(r_q, M_q): the own-chunk argmin takes it; only COUNTED (g_i times the count at the end).
// Two compares that yield 1.0 / 0.0 (FSET), two FMAs and two adds per candidate instead of the select chains and
// shared-memory atomics of cum_until_gather_r / cum_until_scatter_l_fast.  Ties (torch.minimum splits them) are COMMON here --
// the (pos, neg) channels are max(0, v) / min(0, v), exactly 0 half of the time -- so they stay in the main path: r_q == M_q
// is position-level; r_q == h_i only counts when r_q <= M_q, also position-level, so it is one compare against a comparand
// that is NaN otherwise.
template <int NT>
PF_DEV void cum_until_adj_stage(const TC& c, const float* lpl, float* cm) {
  float l[8];
  load8(lpl, c.tid, l);
  float m = f_inf();
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (c.p0 + i < c.nlive) m = fminf(m, l[i]);
  cm[c.tid] = m;
}
template <int UA, int UB, int NT>
PF_DEV void cum_until_adj_walk(const TC& c, bool live, const float* lpl, const float* rpl, const float* cm, float* hl, float* hr,
                               float* own, const float (&g)[8], float (&gl_own)[8], float (&gr_own)[8]) {
  static_assert(UA >= 0 && UB >= UA, "bounded until");
  constexpr int CMAX = (UB + 7) / 8;
  const int p0 = c.p0;
  const int kmax = c.interior ? 0x3fffffff : c.nlive - 1 - p0;
  bool active = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) active = active || g[i] != 0.f;
  active = active && live;
  float lo[8], ro[8], h[8], hacc[8], cgt[8], ceq[8];  // per output: candidates of the tail that left / tied (exact small integers)
  int harg[8];
  load8(lpl, c.tid, lo);
  load8(rpl, c.tid, ro);
  plane_zero8(own, c.tid);  // (private to this thread: only the owner touches its chunk of `own`)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = f_inf();
    hacc[i] = 0.f;
    cgt[i] = ceq[i] = 0.f;
    harg[i] = p0 + i;
  }
  if (active) {  // own chunk: triangular (the l-routes' runs settle into `own`, the r-routes stay in registers)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float pm = f_inf(), run = 0.f;
      int arg = p0 + i;
#pragma unroll
      for (int q = i; q < 8; ++q) {
        if (q > kmax) break;
        if (lo[q] < pm) {  // a new prefix minimum (first index wins ties): settle the old one
          if (run != 0.f) own[pidx(arg)] += run;
          run = 0.f;
          pm = lo[q];
          arg = p0 + q;
        }
        if (q - i >= UA && q - i <= UB) {
          const float tr = cum_route_r(ro[q], pm, g[i]);
          gr_own[q] += tr;
          run += g[i] - tr;
        }
      }
      h[i] = pm;
      harg[i] = arg;
      hacc[i] = run;
    }
  }
  int cend = CMAX;
  if (cend > NT - 1 - c.tid) cend = NT - 1 - c.tid;
  if (cend > (kmax >> 3)) cend = kmax >> 3;
  if (!active) cend = 0;
  float cmin[CMAX + 1];  // cmin[j] = minimum of l over chunks tid + 1 .. tid + j
  cmin[0] = f_inf();
#pragma unroll 1
  for (int cc = 1; cc <= cend; ++cc) cmin[cc] = fminf(cmin[cc - 1], cm[c.tid + cc]);
  float accM = 0.f;  // cotangent routed to the tail's current argmin, not yet settled
  auto chunk = [&](int cc, auto all, auto clip, float (&zl)[8], float (&zr)[8]) {
    float l[8], r[8], Lm[8];
    load8(lpl, c.tid + cc, l);
    load8(rpl, c.tid + cc, r);
    const float m0 = cmin[cc - 1];
    {
      float m = m0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        m = fminf(m, l[k]);
        Lm[k] = m;
      }
    }
#pragma unroll
    for (int k = 7; k >= 0; --k) {
      float rs = 0.f;
      if (!decltype(clip)::value || 8 * cc + k <= kmax) {
        const float M = Lm[k], rq = r[k];
        const float thr = fminf(rq, M);
        // a tie r_q == h_i only counts when the own-chunk argmin holds the minimum (h_i <= M_q), i.e. when r_q <= M_q: a
        // position-level condition, folded into the comparand (NaN never compares equal)
        const float xr = rq <= M ? rq : f_nan();
        float sA = 0.f, sT = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (decltype(all)::value || (8 * cc + k - i >= UA && 8 * cc + k - i <= UB)) {
            const float a = thr < h[i] ? 1.f : 0.f;  // the cotangent leaves the thread (FSET)
            const float e = xr == h[i] ? 1.f : 0.f;  // half of it goes to r_q
            sA = fmaf(a, g[i], sA);
            sT = fmaf(e, g[i], sT);
            cgt[i] += a;
            ceq[i] += e;
          }
        rs = 0.5f * sT;
        const float tr = rq < M ? sA : (rq == M ? 0.5f * sA : 0.f);
        rs += tr;
        accM += sA - tr;
      }
      // a record low of the tail (first index wins ties): everything routed to "the argmin" so far belongs to q
      const bool rec = l[k] < (k ? Lm[k > 0 ? k - 1 : 0] : m0);
      zl[k] = rec ? accM : 0.f;
      accM = rec ? 0.f : accM;
      zr[k] = rs;
    }
  };
  auto step = [&](int cc, auto all) {
    float zl[8], zr[8];
    if (cc <= cend) {
      if (8 * cc + 7 <= kmax) chunk(cc, all, BoolC<false>{}, zl, zr);
      else chunk(cc, all, BoolC<true>{}, zl, zr);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) zl[k] = zr[k] = 0.f;
    }
    if (cc == CMAX) {  // the first step initialises every slot: its writer's numbers, or zeros where there is no writer
      if (c.tid + cc < NT) {
        store8(hl, c.tid + cc, zl);
        store8(hr, c.tid + cc, zr);
      }
      if (c.tid < CMAX) {
        plane_zero8(hl, c.tid);
        plane_zero8(hr, c.tid);
      }
    } else if (c.tid + cc < NT && cc <= cend) {  // one writer per chunk and step: a plain read-modify-write
      float a[8], b[8];
      load8(hl, c.tid + cc, a);
      load8(hr, c.tid + cc, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a[k] += zl[k];
        b[k] += zr[k];
      }
      store8(hl, c.tid + cc, a);
      store8(hr, c.tid + cc, b);
    }
    simt::sync();
  };
  constexpr int CF0 = (UA + 7 + 7) / 8 > 1 ? (UA + 7 + 7) / 8 : 1;
  constexpr int CF1 = UB >= 15 ? (UB - 7) / 8 : 0;
  if constexpr (CF1 >= CF0) {
#pragma unroll
    for (int cc = CMAX; cc > CF1; --cc) step(cc, BoolC<false>{});
#pragma unroll 1
    for (int cc = CF1; cc >= CF0; --cc) step(cc, BoolC<true>{});
#pragma unroll
    for (int cc = CF0 - 1; cc >= 1; --cc) step(cc, BoolC<false>{});
  } else {
#pragma unroll
    for (int cc = CMAX; cc >= 1; --cc) step(cc, BoolC<false>{});
  }
  if (live) {  // what the threads before this one handed over
    float a[8], b[8];
    load8(hl, c.tid, a);
    load8(hr, c.tid, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      gl_own[k] += a[k];
      gr_own[k] += b[k];
    }
  }
  simt::sync();  // (the hand-off planes are reused by the other channel's walk)
  if (!active) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // candidates of output i the tail walk visited: offsets [max(8, i + UA), min(i + UB, kmax, 8 cend + 7)] from p0
    int hi = i + UB;
    if (hi > kmax) hi = kmax;
    if (hi > 8 * cend + 7) hi = 8 * cend + 7;
    const int lo_k = i + UA > 8 ? i + UA : 8;
    const int n = hi - lo_k + 1;
    const float tot = fmaf((float)(n > 0 ? n : 0) - cgt[i] - 0.5f * ceq[i], g[i], hacc[i]);
    if (tot != 0.f) own[pidx(harg[i])] += tot;
  }
  float t[8];
  load8(own, c.tid, t);
#pragma unroll
  for (int k = 0; k < 8; ++k) gl_own[k] += t[k];
}

// ---- gradient rows -> HBM: the thread's 8 rows of D channels, positions inside [p_lo, p_hi)
template <int D>
PF_DEV void store_grad_rows(const TC& c, long long T, const float (&gr)[8][D], float* __restrict__ grad_signal) {
  if (c.p0 + 8 <= c.p_lo || c.p0 >= c.p_hi) return;  // halo thread
  float* dst = grad_signal + ((c.b * T + c.origin + c.p0) * (long long)D);
  if (c.p0 >= c.p_lo && c.p0 + 8 <= c.p_hi && ((u64)dst & 15ull) == 0ull) {
    F4* d4 = reinterpret_cast<F4*>(dst);
#pragma unroll
    for (int k = 0; k < 2 * D; ++k) {
      F4 q;
#pragma unroll
      for (int e = 0; e < 4; ++e) q.v[e] = gr[(4 * k + e) / D][(4 * k + e) % D];
      d4[k] = q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (c.p0 + i >= c.p_lo && c.p0 + i < c.p_hi) {
#pragma unroll
        for (int d = 0; d < D; ++d) dst[i * D + d] = gr[i][d];
      }
  }
}
// D == 2, qualifying launches (TileGeom::use_tma_st): the tile's gradient rows leave through the TMA
// engine.  Each thread parks its 64 bytes in the staging area with the SWIZZLE_64B pattern of
// load_rows_tma (conflict-free 128-bit stores: 4 wavefronts per instruction where the global row
// stores took 16), then one thread issues the tensor store of rows [p_lo, p_hi); the tensor is
// [trace][row][16 floats], so a last, shorter tile is clipped at the end of its trace by the
// hardware.  All threads call it.  stage: 1024-byte aligned, aliases the window scratch.
// The fused mean rides along: its block reduction shares the first barrier, and its atomic is in
// flight while the tile is staged and stored (acc == nullptr: no mean; red: NT / 32 floats).
template <int D, int NT>
PF_DEV void store_grad_rows_tma(const TC& c, const TileGeom& g, bool live, const float (&gr)[8][D], float* __restrict__ grad_signal,
                                float* stage, float part, double* __restrict__ acc, float* red) {
  static_assert(D == 2, "64-byte rows per thread");
  if (!g.use_tma_st) {
    if (live) store_grad_rows<D>(c, g.T, gr, grad_signal);
    if (acc) block_sum_to<NT>(c, part, acc, red);
    return;
  }
  if (acc) {
    float s = part;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += simt::shfl_xor(s, d);
    if ((c.tid & 31) == 0) red[c.tid >> 5] = s;
  }
  simt::sync();  // the last adjoint stage may still be reading the scratch this aliases
  if (acc && c.tid == 32 % NT) {  // (not thread 0: that one issues the tensor store below)
    double tot = 0.0;
    for (int w = 0; w < NT / 32; ++w) tot += (double)red[w];
    simt::atomic_add(acc, tot);
  }
  if (live && c.p0 >= c.p_lo && c.p0 + 8 <= c.p_hi) {
    char* base = reinterpret_cast<char*>(stage) + c.tid * 64;
    const int sw = (c.tid >> 1) & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *reinterpret_cast<F4*>(base + ((k ^ sw) << 4)) = F4{{gr[2 * k][0], gr[2 * k][1], gr[2 * k + 1][0], gr[2 * k + 1][1]}};
  }
  simt::fence_async_smem();
  simt::sync();
  if (c.tid == 0) {
    const int row0 = (int)((c.origin + c.p_lo) >> 3);  // 64-byte row inside the trace: 8 samples of 2 channels
    for (int k = 0; k < g.st_nbox; ++k)
      simt::tma_store_3d(&g.tmap_st, stage + ((c.p_lo >> 3) + k * g.st_rows) * 16, 0, row0 + k * g.st_rows, (int)c.b);
    simt::tma_store_commit_wait();
  }
}
// wide signals (D > 8): one channel at a time
PF_DEV void store_grad_channel(const TC& c, long long T, int D, int ch, const float (&gx)[8], float* __restrict__ grad_signal) {
  float* dst = grad_signal + ((c.b * T + c.origin + c.p0) * (long long)D) + ch;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (c.p0 + i >= c.p_lo && c.p0 + i < c.p_hi) dst[(long long)i * D] = gx[i];
}

}  // namespace fused
}  // namespace pystl
