// TEST INFRASTRUCTURE -- host emulation of the CUDA tile kernel.
//
// Compiles pystl_b200/csrc/tile_eval.h (the single-source CTA body) and
// pystl_b200/csrc/lower.cpp with g++ and runs every "CTA" as a sequential loop
// over its threads (HostExec).  It exists so that the CPU-only test tier can
// check the kernel's lowering, tiling, halo and index arithmetic against the
// oracle without a GPU.  It is built into tests/_emu/ by tests/emu/build.py and
// is never loaded by the pystl_b200 package: the product path is CUDA only.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/pystl_b200.h"
#include "../../pystl_b200/csrc/program.h"
#include "../../pystl_b200/csrc/tile_eval.h"
#include "../../pystl_b200/csrc/tile_bwd.h"

using namespace pystl;

static thread_local char g_err[512];

template <int SEM>
static void run_all(const Lowered& lw, const TileGeom& g, const float* signal, float* out, double* mean_acc) {
  std::vector<float> smem((size_t)g.smem_floats + 64, 0.f);
  // 16-byte align the base like the device's dynamic shared memory
  float* base = smem.data();
  while (reinterpret_cast<uintptr_t>(base) & 15) ++base;
  constexpr int VW = Tr<SEM>::VW;
  for (int64_t b = 0; b < g.B; ++b) {
    for (int j = 0; j < g.n_tiles; ++j) {
      // poison shared memory so that reads of never-written samples show up
      for (int i = 0; i < g.smem_floats; ++i) base[i] = NAN;
      const int64_t e0 = (int64_t)j * g.lout;
      const Ctx c = make_ctx(base, g, lw.prog, e0);
      float* raw = c.raw();
      const float* src = signal + (b * g.T + e0) * g.D;
      for (int i = 0; i < c.nlive * g.D; ++i) raw[i] = src[i];
      HostExec ex(g.nt);
      run_program<SEM>(ex, c, lw.prog, raw);
      store_root<SEM>(ex, c, lw.prog, out + (b * g.T + e0) * VW, g.lout, mean_acc);
    }
  }
}

extern "C" const char* pystl_emu_last_error(void) { return g_err; }

extern "C" int pystl_emu_plan(const pystl_formula_t* f, int64_t T, int32_t D, int32_t backward, pystl_plan_t* plan) {
  Lowered lw;
  int rc = lower_formula(reinterpret_cast<const ::pystl_formula*>(f), backward != 0, &lw, g_err, sizeof(g_err));
  if (rc) return rc;
  TileGeom g;
  rc = plan_geometry(lw, 1, T, D, 0, backward != 0, &g, g_err, sizeof(g_err));
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
  return 0;
}

// force_threads > 0 overrides the CTA size (to exercise multi-tile paths on short traces)
extern "C" int pystl_emu_eval_forward(const pystl_formula_t* f, const float* signal, int64_t B, int64_t T, int32_t D,
                                      float* out, double* mean_acc, int32_t force_threads) {
  Lowered lw;
  int rc = lower_formula(reinterpret_cast<const ::pystl_formula*>(f), false, &lw, g_err, sizeof(g_err));
  if (rc) return rc;
  if (B == 0 || T == 0) return 0;
  TileGeom g;
  rc = plan_geometry(lw, B, T, D, force_threads, false, &g, g_err, sizeof(g_err));
  if (rc) return rc;
  switch (lw.prog.sem) {
    case SEM_CLASSIC: run_all<SEM_CLASSIC>(lw, g, signal, out, mean_acc); break;
    case SEM_SMOOTH: run_all<SEM_SMOOTH>(lw, g, signal, out, mean_acc); break;
    case SEM_CUMULATIVE: run_all<SEM_CUMULATIVE>(lw, g, signal, out, mean_acc); break;
    case SEM_AGM: run_all<SEM_AGM>(lw, g, signal, out, mean_acc); break;
    case SEM_DGMSR: run_all<SEM_DGMSR>(lw, g, signal, out, mean_acc); break;
    default: return PYSTL_E_INVALID;
  }
  return 0;
}

// ---------------------------------------------------------------------------- backward
template <int SEM>
static void run_all_bwd(const Lowered& lw, const TileGeom& g, const float* signal, const float* grad_out, float scale,
                        float* grad_signal) {
  std::vector<float> smem((size_t)g.smem_floats + 64, 0.f);
  float* base = smem.data();
  while (reinterpret_cast<uintptr_t>(base) & 15) ++base;
  for (int64_t b = 0; b < g.B; ++b) {
    for (int j = 0; j < g.n_tiles; ++j) {
      for (int i = 0; i < g.smem_floats; ++i) base[i] = NAN;
      const int64_t t0 = (int64_t)j * g.lout;
      const int64_t e0 = t0 > g.hh ? t0 - g.hh : 0;
      const Ctx c = make_ctx(base, g, lw.prog, e0);
      const int p_lo = (int)(t0 - e0);
      int p_hi = p_lo + g.lout;
      if (p_hi > c.nlive) p_hi = c.nlive;
      float* raw = c.raw();
      const float* src = signal + (b * g.T + e0) * g.D;
      for (int i = 0; i < c.nlive * g.D; ++i) raw[i] = src[i];
      HostExec ex(g.nt);
      run_backward<SEM>(ex, c, lw.prog, raw, grad_out ? grad_out + (b * g.T + e0) : nullptr, scale,
                        grad_signal + (b * g.T + e0) * g.D, p_lo, p_hi);
    }
  }
}

extern "C" int pystl_emu_eval_backward(const pystl_formula_t* f, const float* signal, int64_t B, int64_t T, int32_t D,
                                       const float* grad_out, float scale, float* grad_signal, int32_t force_threads) {
  Lowered lw;
  int rc = lower_formula(reinterpret_cast<const ::pystl_formula*>(f), true, &lw, g_err, sizeof(g_err));
  if (rc) return rc;
  if (B == 0 || T == 0) return 0;
  TileGeom g;
  rc = plan_geometry(lw, B, T, D, force_threads, true, &g, g_err, sizeof(g_err));
  if (rc) return rc;
  switch (lw.prog.sem) {
    case SEM_CLASSIC: run_all_bwd<SEM_CLASSIC>(lw, g, signal, grad_out, scale, grad_signal); break;
    case SEM_SMOOTH: run_all_bwd<SEM_SMOOTH>(lw, g, signal, grad_out, scale, grad_signal); break;
    case SEM_AGM: run_all_bwd<SEM_AGM>(lw, g, signal, grad_out, scale, grad_signal); break;
    case SEM_DGMSR: run_all_bwd<SEM_DGMSR>(lw, g, signal, grad_out, scale, grad_signal); break;
    default: snprintf(g_err, sizeof(g_err), "no backward for this semantics"); return PYSTL_E_UNSUPPORTED;
  }
  return 0;
}

// ---------------------------------------------------------------------------- fused tier
// Lower (SSA) + plan + generate the fused kernel source; the test harness compiles that source
// with g++ against simt_host.h (tests/emu/fused_runner.h) and runs it on the CPU.
#include "fused_host.h"
extern "C" int pystl_emu_fused_sizes(int32_t* prog_bytes, int32_t* geom_bytes) {
  *prog_bytes = (int32_t)sizeof(TileProgram);
  *geom_bytes = (int32_t)sizeof(TileGeom);
  return 0;
}
extern "C" int pystl_emu_fused_prepare(const pystl_formula_t* f, int64_t B, int64_t T, int32_t D, int32_t backward,
                                       int32_t force_nt, char* src, int32_t src_len, void* prog, void* geom,
                                       int32_t* smem_floats) {
  Lowered lw;
  int rc = lower_formula_ex(reinterpret_cast<const ::pystl_formula*>(f), backward != 0, true, &lw, g_err, sizeof(g_err));
  if (rc) return rc;
  TileGeom g;
  rc = fused_geometry(lw, B, T, D, force_nt, backward != 0, &g, g_err, sizeof(g_err));
  if (rc) return rc;
  FusedSource fs;
  rc = fused_generate(lw, D, g.nt, 1, backward != 0, &fs, g_err, sizeof(g_err));
  if (rc) return rc;
  if ((int)fs.source.size() + 1 > src_len) {
    snprintf(g_err, sizeof(g_err), "source buffer too small");
    return PYSTL_E_INVALID;
  }
  memcpy(src, fs.source.c_str(), fs.source.size() + 1);
  memcpy(prog, &lw.prog, sizeof(TileProgram));
  memcpy(geom, &g, sizeof(TileGeom));
  *smem_floats = fs.smem_floats;
  return 0;
}
