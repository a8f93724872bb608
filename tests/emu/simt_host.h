// Host stand-in for the CUDA execution model (TEST INFRASTRUCTURE, never part of the product).
//
// The fused tile kernels (pystl_b200/csrc/fused_rt.cuh + the generated glue) are written in
// plain SIMT style.  To check their index arithmetic, halos and barrier placement against the
// oracle without a GPU, the CPU test tier compiles the very same source with g++ against this
// header: every CUDA thread of a CTA becomes a ucontext fiber with its own stack, and
// __syncthreads / warp collectives are cooperative yields of a round-robin scheduler.
// Execution is deterministic (thread 0 runs to its next barrier, then thread 1, ...).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <vector>

namespace simt {

struct Launch {
  int nt = 0;
  int bid = 0;
  int bid2 = 0;
  int cur = 0;  // fiber being run
  int alive = 0;
  unsigned long long bar_gen = 0;
  int bar_count = 0;
  std::vector<unsigned long long> wbar_gen;  // per-warp barriers (warp collectives)
  std::vector<int> wbar_count;
  std::vector<float> wslot;  // one exchange slot per thread
  std::vector<char> done;
  std::vector<ucontext_t> ctx;
  std::vector<char*> stacks;
  ucontext_t main_ctx;
  float* dyn_smem = nullptr;
  void (*body)(void*) = nullptr;
  void* arg = nullptr;
};

inline Launch*& cur_launch() {
  static thread_local Launch* l = nullptr;
  return l;
}

inline int tid() { return cur_launch()->cur; }
inline int bid_x() { return cur_launch()->bid; }
inline int bid_y() { return cur_launch()->bid2; }
inline int nthreads() { return cur_launch()->nt; }
inline float* dyn_smem() { return cur_launch()->dyn_smem; }

// switch to the next fiber that has not finished (or back to the launcher when none is left)
inline void yield() {
  Launch* L = cur_launch();
  const int me = L->cur;
  int nxt = me;
  for (int k = 0; k < L->nt; ++k) {
    nxt = (nxt + 1) % L->nt;
    if (!L->done[nxt]) break;
  }
  if (nxt == me) return;
  L->cur = nxt;
  swapcontext(&L->ctx[me], &L->ctx[nxt]);
}

inline void sync() {  // __syncthreads
  Launch* L = cur_launch();
  const unsigned long long gen = L->bar_gen;
  if (++L->bar_count == L->alive) {
    L->bar_count = 0;
    ++L->bar_gen;
  } else {
    while (L->bar_gen == gen) yield();
  }
}

inline void warp_sync() {
  Launch* L = cur_launch();
  const int w = L->cur >> 5;
  const int lanes = (L->nt - 32 * w) < 32 ? (L->nt - 32 * w) : 32;
  const unsigned long long gen = L->wbar_gen[w];
  if (++L->wbar_count[w] == lanes) {
    L->wbar_count[w] = 0;
    ++L->wbar_gen[w];
  } else {
    while (L->wbar_gen[w] == gen) yield();
  }
}

inline void syncwarp() { warp_sync(); }

// warp collectives (all lanes of the warp must call them)
inline float shfl_idx(float v, int src_lane) {
  Launch* L = cur_launch();
  const int base = L->cur & ~31;
  L->wslot[L->cur] = v;
  warp_sync();
  const int s = base + (src_lane & 31);
  const float r = s < L->nt ? L->wslot[s] : v;
  warp_sync();
  return r;
}
inline float shfl_down(float v, int delta) {
  const int lane = tid() & 31;
  return shfl_idx(v, lane + delta < 32 ? lane + delta : lane);
}
inline float shfl_xor(float v, int mask) { return shfl_idx(v, (tid() & 31) ^ mask); }
inline int redux_max_i(int v) {
  Launch* L = cur_launch();
  const int base = L->cur & ~31;
  float f;
  memcpy(&f, &v, 4);
  L->wslot[L->cur] = f;
  warp_sync();
  int m = v;
  for (int l = 0; l < 32 && base + l < L->nt; ++l) {
    int o;
    memcpy(&o, &L->wslot[base + l], 4);
    if (o > m) m = o;
  }
  warp_sync();
  return m;
}

inline void atomic_add(float* p, float v) { *p += v; }
inline void atomic_add(double* p, double v) { *p += v; }
inline void fence_async_smem() {}
// (TileGeom::use_tma is never set on the host: the TMA tile load is a GPU-only path)
inline void mbar_init(void*) { abort(); }
inline void mbar_expect_tx(void*, unsigned) { abort(); }
inline void mbar_wait(void*, unsigned) { abort(); }
inline void tma_load_2d(void*, const void*, int, int, void*) { abort(); }
inline void tma_store_3d(const void*, const void*, int, int, int) { abort(); }
inline void tma_store_commit_wait() { abort(); }
inline void bulk_store_and_wait(void* dst_global, const void* src_smem, unsigned bytes) { memcpy(dst_global, src_smem, bytes); }

inline void fiber_entry() {
  Launch* L = cur_launch();
  L->body(L->arg);
  L->done[L->cur] = 1;
  --L->alive;
  // a barrier the others wait on may now be complete
  if (L->alive > 0 && L->bar_count == L->alive) {
    L->bar_count = 0;
    ++L->bar_gen;
  }
  if (L->alive == 0) {
    setcontext(&L->main_ctx);
  }
  const int me = L->cur;
  int nxt = me;
  for (int k = 0; k < L->nt; ++k) {
    nxt = (nxt + 1) % L->nt;
    if (!L->done[nxt]) break;
  }
  L->cur = nxt;
  setcontext(&L->ctx[nxt]);
}

// run one CTA: `body(arg)` is executed by nt fibers
inline void run_cta(int bid_, int bid_y_, int nt, size_t smem_bytes, void (*body)(void*), void* arg) {
  static thread_local Launch L;
  static const size_t kStack = 256 * 1024;
  if ((int)L.stacks.size() < nt) {
    const size_t old = L.stacks.size();
    L.stacks.resize(nt);
    for (size_t i = old; i < (size_t)nt; ++i) L.stacks[i] = (char*)malloc(kStack);
  }
  L.nt = nt;
  L.bid = bid_;
  L.bid2 = bid_y_;
  L.cur = 0;
  L.alive = nt;
  L.bar_gen = 0;
  L.bar_count = 0;
  L.wbar_gen.assign((nt + 31) / 32, 0);
  L.wbar_count.assign((nt + 31) / 32, 0);
  L.wslot.assign(nt, 0.f);
  L.done.assign(nt, 0);
  L.ctx.resize(nt);
  L.body = body;
  L.arg = arg;
  void* mem = nullptr;
  if (posix_memalign(&mem, 128, smem_bytes + 128)) abort();
  // poison: reads of shared memory nobody wrote show up as NaN
  memset(mem, 0xff, smem_bytes + 128);
  L.dyn_smem = (float*)mem;
  cur_launch() = &L;
  for (int i = 0; i < nt; ++i) {
    getcontext(&L.ctx[i]);
    L.ctx[i].uc_stack.ss_sp = L.stacks[i];
    L.ctx[i].uc_stack.ss_size = kStack;
    L.ctx[i].uc_link = nullptr;
    makecontext(&L.ctx[i], (void (*)())fiber_entry, 0);
  }
  swapcontext(&L.main_ctx, &L.ctx[0]);
  free(mem);
  cur_launch() = nullptr;
}

}  // namespace simt
