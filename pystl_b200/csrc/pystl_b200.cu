// libpystl_b200.so -- the C ABI of include/pystl_b200.h (host side).
//
// One CTA evaluates the whole formula over one tile of one trace
// (tile_eval.h); the signal tile is staged in shared memory by a TMA bulk copy
// (cp.async.bulk, falling back to plain loads when the tile is not 16-byte
// aligned), every intermediate node value stays in shared memory, only the
// root robustness goes back to HBM.  Bandwidth-bound windowed reductions: no
// tensor cores on purpose.  Kernels: kernels.cuh, instantiated per semantics in
// kernels_sem.cu.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/pystl_b200.h"
#include "fused_host.h"
#include "program.h"

namespace pystl {
template <int SEM>
int launch_fwd_sem(const Lowered& lw, const TileGeom& g, const float* signal, float* out, double* mean_acc,
                   cudaStream_t stream, char* err, int err_len);
template <int SEM>
int launch_bwd_sem(const Lowered& lw, const TileGeom& g, const float* signal, const float* grad_out, float scale,
                   float* grad_signal, cudaStream_t stream, char* err, int err_len);
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

static int prepare(const pystl_formula_t* f, int64_t B, int64_t T, int32_t D, bool backward, Lowered* lw, TileGeom* g) {
  int rc = lower_formula(reinterpret_cast<const ::pystl_formula*>(f), backward, lw, g_err, sizeof(g_err));
  if (rc) return rc;
  return plan_geometry(*lw, B, T, D, forced_threads(), backward, g, g_err, sizeof(g_err));
}

// ---------------------------------------------------------------------- fused tier
// PYSTL_B200_TIER = auto (default: fused when the formula is supported, else interpreter),
// interp, or fused (fail instead of falling back; tests)
enum { TIER_AUTO = 0, TIER_INTERP = 1, TIER_FUSED = 2 };
static int tier_mode() {
  const char* e = getenv("PYSTL_B200_TIER");
  if (!e || !*e) return TIER_AUTO;
  if (!strcmp(e, "interp")) return TIER_INTERP;
  if (!strcmp(e, "fused")) return TIER_FUSED;
  return TIER_AUTO;
}

// registers per thread the fused kernels are compiled for: min CTAs per SM by CTA size
static int fused_minb(int nt, bool backward, const TileProgram& prog) {
  const int sem = prog.sem;
  bool has_until = false;  // the O(b) walks keep 8 running states live: they want their registers
  for (int i = 0; i < prog.n_instr; ++i) has_until |= prog.ins[i].op == T_UNTIL;
  const char* e = getenv("PYSTL_B200_FUSED_MINB");
  if (e && *e) return atoi(e) > 0 ? atoi(e) : 1;
  // The backward kernels keep the forward values AND the adjoints of 8 samples in registers, but
  // they are latency bound: more resident warps beat fewer spills.  Register targets measured on
  // B200 (profiles/README.md, sweep of CTA size x CTAs per SM on cfg2 / cfg3): smooth 80
  // (1.81 -> 1.77 ms), agm 72 (3.20 -> 2.86 ms), dgmsr 56 (4.05 -> 3.55 ms).
  int bwd_regs = 96;
  if (sem == SEM_SMOOTH) bwd_regs = (nt == 288 && !has_until) ? 72 : 80;  // (the wide tile of codegen.cpp: 3 CTAs of 9 warps, operands re-lifted)
  else if (sem == SEM_AGM) bwd_regs = 72;
  else if (sem == SEM_DGMSR) bwd_regs = 56;
  // the lockstep Until adjoints keep two sets of 8 running states and the hand-off numbers live: at 80 registers the
  // spills reach DRAM (ncu: 1.6 x the algorithmic traffic); 128 registers, cfg5 Smooth value + gradient 10.0 -> 8.7 ms
  if (has_until && (sem == SEM_SMOOTH || sem == SEM_AGM || sem == SEM_DGMSR)) bwd_regs = 128;
  const char* r = getenv(backward ? "PYSTL_B200_FUSED_REGS_BWD" : "PYSTL_B200_FUSED_REGS");
  // forward kernels: 56 registers (cfg2 0.785 -> 0.773 ms), 48 for agm / dgmsr (cfg3 1.05 -> 0.97 ms)
  const int fwd_regs = has_until ? 64 : ((sem == SEM_AGM || sem == SEM_DGMSR) ? 48 : 56);
  const int target_regs = (r && atoi(r) >= 32) ? atoi(r) : (backward ? bwd_regs : fwd_regs);
  int minb = 65536 / (nt * target_regs);
  return minb < 1 ? 1 : (minb > 8 ? 8 : minb);
}

static int fused_prepare(const pystl_formula_t* f, int64_t B, int64_t T, int32_t D, bool backward, Lowered* lw, TileGeom* g,
                         FusedSource* fs) {
  int rc = lower_formula_ex(reinterpret_cast<const ::pystl_formula*>(f), backward, true, lw, g_err, sizeof(g_err));
  if (rc) return rc;
  rc = fused_geometry(*lw, B, T, D, forced_threads(), backward, g, g_err, sizeof(g_err));
  if (rc) return rc;
  const int minb = fused_minb(g->nt, backward, lw->prog);
  rc = fused_generate(*lw, D, g->nt, minb, backward, fs, g_err, sizeof(g_err));
  if (rc) return rc;
  if (fs->smem_floats * 4 > 227 * 1024) return set_err(PYSTL_E_UNSUPPORTED, "fused kernel needs %d bytes of shared memory", fs->smem_floats * 4);
  {
    // shared memory may admit fewer CTAs per SM than the register target assumes: give the registers back
    const int by_smem = std::max(1, 228 * 1024 / (fs->smem_floats * 4 + 1024));
    if (by_smem < minb && !getenv("PYSTL_B200_FUSED_MINB")) {
      rc = fused_generate(*lw, D, g->nt, by_smem, backward, fs, g_err, sizeof(g_err));
      if (rc) return rc;
    }
  }
  g->smem_floats = fs->smem_floats;
  // L2 prefetch distance in traces (fused_rt.cuh: prefetch_rows): a little more than one wave of
  // resident CTAs ahead.  TileGeom::scr_planes is unused by the fused tier and carries it.
  {
    const char* e = getenv("PYSTL_B200_PREFETCH");
    const int regs = backward ? 96 : 64;
    int per_sm = std::min(65536 / (g->nt * regs), std::min(227 * 1024 / (fs->smem_floats * 4 + 1024), 2048 / g->nt));
    if (per_sm < 1) per_sm = 1;
    g->scr_planes = e ? atoi(e) : (148 * per_sm * 5) / 4;
  }
  return PYSTL_OK;
}

// ---------------------------------------------------------------------- prepared launches
// Lowering a formula, choosing the geometry, generating the kernel source and hashing it cost tens of
// microseconds per call -- more than the kernel of a small launch (cfg1: 29 us).  A call is therefore
// keyed by the BYTES of everything that determines the launch except the buffers and the batch size:
// the formula (structure and numerics), T, D, the pass, the device and the experiment switches.  The
// entry holds the lowered program, the geometry, the loaded kernel and the last tensor maps that were
// encoded, so a steady-state call is one hash lookup and one cudaLaunchKernel (and nothing in it
// synchronises or allocates: the path can be captured into a CUDA graph).
struct Prepared {
  int rc = PYSTL_OK;  // PYSTL_E_UNSUPPORTED: this formula runs on the interpreter tier
  std::string err;
  Lowered lw;
  TileGeom g;  // B = 1, no tensor maps
  std::string name;
  int smem_floats = 0;
  FusedSource fs;        // kept until the kernel is loaded
  FusedKernel fk;        // loaded on the first launch (planning and precompiling need no GPU)
  bool loaded = false;
  std::mutex mu;         // guards fk / loaded and the tensor-map memos
  const void* ld_ptr = nullptr;
  int64_t ld_B = -1;
  int ld_use = 0;
  unsigned long long ld_map[16];
  const void* st_ptr = nullptr;
  int64_t st_B = -1;
  int st_use = 0, st_rows = 0, st_nbox = 0;
  unsigned long long st_map[16];
};
static std::mutex g_prep_mu;
static std::unordered_map<std::string, std::shared_ptr<Prepared>> g_prep;
static const char* const kKnobs[] = {"PYSTL_B200_FORCE_THREADS", "PYSTL_B200_FUSED_MINB", "PYSTL_B200_FUSED_REGS", "PYSTL_B200_FUSED_REGS_BWD",
                                     "PYSTL_B200_PREFETCH",      "PYSTL_B200_MERGE",      "PYSTL_B200_NVRTC_FLAGS", "PYSTL_B200_TMA",
                                     "PYSTL_B200_TMA_STORE",     "PYSTL_B200_CACHE",      "PYSTL_B200_RELIFT",
                                     "PYSTL_B200_UNTIL_EXP"};

template <class X>
static void key_put(std::string* k, const X* p, size_t n) {
  if (p && n) k->append(reinterpret_cast<const char*>(p), n * sizeof(X));
  k->push_back('|');
}
static bool prepared_key(const pystl_formula_t* f, int64_t T, int32_t D, bool backward, std::string* k) {
  if (!f || f->n_nodes < 0 || f->n_children < 0 || f->n_preds < 0 || f->n_weights < 0) return false;
  if ((f->n_nodes && !f->nodes) || (f->n_children && !f->children) || (f->n_preds && !f->preds) || (f->n_weights && !f->weights)) return false;
  k->reserve(512);
  int dev = 0;
  cudaGetDevice(&dev);
  const int64_t head[8] = {f->semantics, f->p, f->n_nodes, f->n_children, T, D, backward ? 1 : 0, dev};
  key_put(k, head, 8);
  key_put(k, &f->temperature, 1);
  key_put(k, &f->eps, 1);
  key_put(k, f->nodes, (size_t)f->n_nodes);
  key_put(k, f->children, (size_t)f->n_children);
  key_put(k, f->preds, (size_t)f->n_preds);
  key_put(k, f->weights, (size_t)f->n_weights);
  for (const char* name : kKnobs) {
    const char* e = getenv(name);
    if (e) k->append(e);
    k->push_back('|');
  }
  return true;
}

// the prepared launch of (formula, T, D, pass); built on a miss.  Never null.
static std::shared_ptr<Prepared> fused_prepared(const pystl_formula_t* f, int64_t T, int32_t D, bool backward) {
  std::string key;
  const bool cacheable = prepared_key(f, T, D, backward, &key);
  if (cacheable) {
    std::lock_guard<std::mutex> lk(g_prep_mu);
    auto it = g_prep.find(key);
    if (it != g_prep.end()) return it->second;
  }
  auto p = std::make_shared<Prepared>();
  p->rc = fused_prepare(f, 1, T, D, backward, &p->lw, &p->g, &p->fs);
  if (p->rc) {
    p->err = g_err;
    p->fs = FusedSource();
  } else {
    p->name = p->fs.name;
    p->smem_floats = p->fs.smem_floats;
  }
  if (cacheable) {
    std::lock_guard<std::mutex> lk(g_prep_mu);
    // (a training loop that moves a threshold makes a new key every step: bound the table)
    if (g_prep.size() >= 1024) g_prep.clear();
    g_prep[key] = p;
  }
  return p;
}

// D == 2: describe the signal as rows of 64 bytes for the TMA tile load of the fused kernels
// (fused_rt.cuh: load_rows_tma).  The launch qualifies when every tile starts on a 64-byte row
// (tile origins are multiples of 8 samples = 64 bytes; the trace stride T * D * 4 must be one too),
// the box of one tile has at most 256 rows and the driver offers cuTensorMapEncodeTiled.
// PYSTL_B200_TMA=0 switches the path off.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
static void fused_tma_setup(TileGeom* g, const float* signal, int sem) {
  g->use_tma = 0;
  const char* e = getenv("PYSTL_B200_TMA");  // 0: off, 2: on wherever the shape qualifies (tests)
  if (e && atoi(e) == 0) return;
  const bool force = e && atoi(e) == 2;
  // (measured: +17 % on cfg1x, +5 % on cfg2 forward; the agm / dgmsr kernels are latency bound and
  // do not gain from it: 2.93 vs 2.95 ms on cfg3)
  if (!force && (sem == SEM_AGM || sem == SEM_DGMSR)) return;
  // (value-and-gradient kernels other than the classic one: 1.646 ms without, 1.671 ms with it on cfg2)
  if (!force && g->bwd && sem != SEM_CLASSIC) return;
  // a launch of a few waves is latency bound: the barrier + TMA round trip at the head of every CTA
  // costs more than the wavefronts it saves (cfg1, 1024 CTAs: 47 us without, 70 us with it)
  if (!force && g->B * (int64_t)g->n_tiles < 8192) return;
  static_assert(sizeof(CUtensorMap) == sizeof(g->tmap), "CUtensorMap is 128 bytes");
  if (g->D != 2 && g->D != 4 && g->D != 8) return;
  const int rowb = g->D == 2 ? 64 : 128;                 // bytes per tensor row (fused_rt.cuh: load_rows_tma)
  const int trows = g->nt * (32 * g->D / rowb);          // rows of one tile
  const int nbox = (trows + 255) / 256;                  // boxes of at most 256 rows
  if (trows % nbox != 0 || g->B <= 0 || g->T <= 0) return;
  if ((g->T * g->D * 4) % rowb != 0 || (reinterpret_cast<uintptr_t>(signal) & (uintptr_t)(rowb - 1)) != 0) return;
  const uint64_t rows = (uint64_t)g->B * (uint64_t)g->T * (uint64_t)g->D * 4u / (uint64_t)rowb;
  if (rows > 0x7fffffffull) return;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return;
  const cuuint64_t dims[2] = {(cuuint64_t)(rowb / 4), rows};
  const cuuint64_t strides[1] = {(cuuint64_t)rowb};
  const cuuint32_t box[2] = {(cuuint32_t)(rowb / 4), (cuuint32_t)(trows / nbox)};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(signal), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          g->D == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return;
  memcpy(g->tmap, &tm, sizeof(tm));
  g->use_tma = 1;
}

// ... and grad_signal as [trace][64-byte row][16 floats] for the TMA store of the gradient rows
// (fused_rt.cuh: store_grad_rows_tma): boxes of at most 256 rows, an even number of rows each.
static void fused_tma_setup_store(TileGeom* g, float* grad_signal, int sem) {
  g->use_tma_st = 0;
  const char* e = getenv("PYSTL_B200_TMA_STORE");
  if (e && atoi(e) == 0) return;
  if (sem == SEM_CLASSIC) return;  // its narrow-signal gradient leaves as a ready-made tile image already
  if (sem == SEM_DGMSR) return;    // measured: cfg2 1.72 -> 1.65 ms, cfg3 agm 2.91 -> 2.80 ms, but dgmsr 3.58 -> 3.62 ms
  if (g->D != 2 || g->B <= 0 || g->T <= 0 || (g->T * g->D) % 16 != 0 || (reinterpret_cast<uintptr_t>(grad_signal) & 63u) != 0) return;
  if (g->hh % 16 != 0 || g->lout % 16 != 0) return;
  const int rows = g->lout / 8, nbox = (rows + 255) / 256;
  if (rows % nbox != 0 || ((rows / nbox) & 1) != 0 || g->B > 0x7fffffffLL) return;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return;
  const cuuint64_t dims[3] = {16, (cuuint64_t)(g->T * g->D / 16), (cuuint64_t)g->B};
  const cuuint64_t strides[2] = {64, (cuuint64_t)(g->T * g->D * 4)};
  const cuuint32_t box[3] = {16, (cuuint32_t)(rows / nbox), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap tm;
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, grad_signal, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return;
  memcpy(g->tmap_st, &tm, sizeof(tm));
  g->st_rows = rows / nbox;
  g->st_nbox = nbox;
  g->use_tma_st = 1;
}

static int fused_launch(Prepared& p, const TileGeom& g, void** args, cudaStream_t stream) {
  void* kernel = nullptr;
  {
    std::lock_guard<std::mutex> lk(p.mu);
    if (!p.loaded) {
      int rc = fused_jit_get(p.fs, &p.fk, g_err, sizeof(g_err));
      if (rc) return rc;
      p.loaded = true;
      p.fs.source.clear();
      p.fs.source.shrink_to_fit();
    }
    kernel = p.fk.kernel;
  }
  if (g.B > 0x7fffffffLL || g.n_tiles > 65535) return set_err(PYSTL_E_UNSUPPORTED, "grid of %lld x %d tiles is too large", (long long)g.B, g.n_tiles);
  cudaError_t e = cudaLaunchKernel(reinterpret_cast<const void*>(kernel), dim3((unsigned)g.B, (unsigned)g.n_tiles), dim3((unsigned)g.nt), args,
                                   (size_t)p.smem_floats * 4, stream);
  if (e != cudaSuccess) return set_err(PYSTL_E_CUDA, "fused kernel launch: %s", cudaGetErrorString(e));
  return PYSTL_OK;
}

// this call's geometry: the prepared one with the batch size and the tensor maps of these buffers
// (the maps are re-encoded only when a buffer or the batch size changes)
static void fused_bind(Prepared& p, int64_t B, const float* signal, float* grad_signal, TileGeom* g) {
  *g = p.g;
  g->B = B;
  std::lock_guard<std::mutex> lk(p.mu);
  if (p.ld_ptr != signal || p.ld_B != B) {
    fused_tma_setup(g, signal, p.lw.prog.sem);
    p.ld_ptr = signal;
    p.ld_B = B;
    p.ld_use = g->use_tma;
    memcpy(p.ld_map, g->tmap, sizeof(p.ld_map));
  } else {
    g->use_tma = p.ld_use;
    if (p.ld_use) memcpy(g->tmap, p.ld_map, sizeof(p.ld_map));
  }
  if (!grad_signal) return;
  if (p.st_ptr != grad_signal || p.st_B != B) {
    fused_tma_setup_store(g, grad_signal, p.lw.prog.sem);
    p.st_ptr = grad_signal;
    p.st_B = B;
    p.st_use = g->use_tma_st;
    p.st_rows = g->st_rows;
    p.st_nbox = g->st_nbox;
    memcpy(p.st_map, g->tmap_st, sizeof(p.st_map));
  } else {
    g->use_tma_st = p.st_use;
    g->st_rows = p.st_rows;
    g->st_nbox = p.st_nbox;
    if (p.st_use) memcpy(g->tmap_st, p.st_map, sizeof(p.st_map));
  }
}

// Compile the fused kernel of (formula, T, D) into the on-disk cache; needs no GPU.
// Returns PYSTL_E_UNSUPPORTED when the formula runs on the interpreter tier.
extern "C" int pystl_b200_fused_precompile(const pystl_formula_t* f, int64_t T, int32_t D, int32_t backward) {
  Lowered lw;
  TileGeom g;
  FusedSource fs;
  int rc = fused_prepare(f, 1, T, D, backward != 0, &lw, &g, &fs);
  if (rc) return rc;
  return fused_jit_precompile(fs, g_err, sizeof(g_err));
}

// Geometry the fused tier uses for (formula, T, D); PYSTL_E_UNSUPPORTED: interpreter tier.
extern "C" int pystl_b200_fused_plan(const pystl_formula_t* f, int64_t T, int32_t D, int32_t backward, pystl_plan_t* plan) {
  if (!plan) return set_err(PYSTL_E_INVALID, "plan is NULL");
  Lowered lw;
  TileGeom g;
  FusedSource fs;
  int rc = fused_prepare(f, 1, T, D, backward != 0, &lw, &g, &fs);
  if (rc) return rc;
  plan->out_channels = lw.out_channels;
  plan->lookahead = lw.lookahead;
  plan->start_slack = lw.prog.start_slack;
  plan->raises_always = lw.prog.raises_always;
  plan->threads = g.nt;
  plan->tile_ext = g.lext;
  plan->tile_out = g.lout;
  plan->n_tiles = g.n_tiles;
  plan->smem_bytes = fs.smem_floats * 4;
  plan->n_instr = lw.prog.n_instr;
  plan->n_planes = lw.prog.n_slots;
  return PYSTL_OK;
}

// The generated source (debugging / inspection).  Returns its length, or a negative status.
extern "C" int pystl_b200_fused_source(const pystl_formula_t* f, int64_t T, int32_t D, int32_t backward, char* buf, int32_t len) {
  Lowered lw;
  TileGeom g;
  FusedSource fs;
  int rc = fused_prepare(f, 1, T, D, backward != 0, &lw, &g, &fs);
  if (rc) return rc;
  if (buf && len > 0) snprintf(buf, (size_t)len, "%s", fs.source.c_str());
  return (int)fs.source.size();
}

extern "C" int pystl_b200_abi_version(void) { return PYSTL_B200_ABI_VERSION; }
extern "C" const char* pystl_b200_last_error(void) { return g_err; }

// open-ended windows beyond one tile: scan.cu
namespace pystl {
int open_window_forward(int32_t semantics, int32_t mode, int32_t a, float temperature, float eps, int32_t p, const float* y,
                        int64_t B, int64_t T, float* out, int32_t* arg, void* stream, char* err, size_t err_cap);
int open_window_backward(int32_t semantics, int32_t mode, int32_t a, float temperature, float eps, int32_t p, const float* y,
                         const float* out, const int32_t* arg, const float* grad_out, int64_t B, int64_t T, float* grad_y,
                         void* stream, char* err, size_t err_cap);
int open_until_forward(int32_t semantics, int32_t a, const float* l, const float* r, int64_t B, int64_t T, float* out, int32_t* key,
                       void* stream, char* err, size_t err_cap);
int open_until_backward(int32_t semantics, int32_t a, const int32_t* key, const float* grad_out, int64_t B, int64_t T, float* grad_l,
                        float* grad_r, void* stream, char* err, size_t err_cap);
int assemble_channels(const float* signal, int64_t rows, int32_t D, const int32_t* channels, int32_t n_channels, const float* const* cols,
                      int32_t n_cols, float* out, void* stream, char* err, size_t err_cap);
int scatter_channels(const float* grad, int64_t rows, int32_t D, const int32_t* channels, int32_t n_channels, float* grad_signal,
                     float* const* grad_cols, int32_t n_cols, void* stream, char* err, size_t err_cap);
// CT-STL on value traces: select.cu
int ctstl_cumulative_forward(const float* y, int64_t B, int64_t T, int32_t a, int32_t b_end, int32_t k, float* out, int32_t* arg,
                             void* stream, char* err, size_t err_cap);
int ctstl_cumulative_backward(const int32_t* arg, const float* grad_out, int64_t B, int64_t T, int32_t a, int32_t b_end, float* grad_y,
                              void* stream, char* err, size_t err_cap);
int ctstl_until_forward(const float* l, const float* r, int64_t B, int64_t T, int32_t a, int32_t b_end, float* out, int32_t* key,
                        void* stream, char* err, size_t err_cap);
int ctstl_until_backward(const int32_t* key, const float* grad_out, int64_t B, int64_t T, int32_t a, int32_t b_end, float* grad_l,
                         float* grad_r, void* stream, char* err, size_t err_cap);
}  // namespace pystl
extern "C" int pystl_b200_ctstl_cumulative_forward(const float* y, int64_t B, int64_t T, int32_t a, int32_t b_end, int32_t k,
                                                   float* out, int32_t* arg, void* stream) {
  return pystl::ctstl_cumulative_forward(y, B, T, a, b_end, k, out, arg, stream, g_err, sizeof(g_err));
}
extern "C" int pystl_b200_ctstl_cumulative_backward(const int32_t* arg, const float* grad_out, int64_t B, int64_t T, int32_t a,
                                                    int32_t b_end, float* grad_y, void* stream) {
  return pystl::ctstl_cumulative_backward(arg, grad_out, B, T, a, b_end, grad_y, stream, g_err, sizeof(g_err));
}
extern "C" int pystl_b200_ctstl_until_forward(const float* l, const float* r, int64_t B, int64_t T, int32_t a, int32_t b_end,
                                              float* out, int32_t* key, void* stream) {
  return pystl::ctstl_until_forward(l, r, B, T, a, b_end, out, key, stream, g_err, sizeof(g_err));
}
extern "C" int pystl_b200_ctstl_until_backward(const int32_t* key, const float* grad_out, int64_t B, int64_t T, int32_t a,
                                               int32_t b_end, float* grad_l, float* grad_r, void* stream) {
  return pystl::ctstl_until_backward(key, grad_out, B, T, a, b_end, grad_l, grad_r, stream, g_err, sizeof(g_err));
}
extern "C" int pystl_b200_assemble_channels(const float* signal, int64_t rows, int32_t D, const int32_t* channels, int32_t n_channels,
                                            const float* const* cols, int32_t n_cols, float* out, void* stream) {
  return pystl::assemble_channels(signal, rows, D, channels, n_channels, cols, n_cols, out, stream, g_err, sizeof(g_err));
}
extern "C" int pystl_b200_scatter_channels(const float* grad, int64_t rows, int32_t D, const int32_t* channels, int32_t n_channels,
                                           float* grad_signal, float* const* grad_cols, int32_t n_cols, void* stream) {
  return pystl::scatter_channels(grad, rows, D, channels, n_channels, grad_signal, grad_cols, n_cols, stream, g_err, sizeof(g_err));
}
extern "C" int pystl_b200_open_until_forward(int32_t semantics, int32_t a, const float* l, const float* r, int64_t B, int64_t T,
                                             float* out, int32_t* key, void* stream) {
  return pystl::open_until_forward(semantics, a, l, r, B, T, out, key, stream, g_err, sizeof(g_err));
}
extern "C" int pystl_b200_open_until_backward(int32_t semantics, int32_t a, const int32_t* key, const float* grad_out, int64_t B,
                                              int64_t T, float* grad_l, float* grad_r, void* stream) {
  return pystl::open_until_backward(semantics, a, key, grad_out, B, T, grad_l, grad_r, stream, g_err, sizeof(g_err));
}
extern "C" int pystl_b200_open_window_forward(int32_t semantics, int32_t mode, int32_t a, float temperature, float eps, int32_t p,
                                              const float* y, int64_t B, int64_t T, float* out, int32_t* arg, void* stream) {
  return pystl::open_window_forward(semantics, mode, a, temperature, eps, p, y, B, T, out, arg, stream, g_err, sizeof(g_err));
}
extern "C" int pystl_b200_open_window_backward(int32_t semantics, int32_t mode, int32_t a, float temperature, float eps, int32_t p,
                                               const float* y, const float* out, const int32_t* arg, const float* grad_out,
                                               int64_t B, int64_t T, float* grad_y, void* stream) {
  return pystl::open_window_backward(semantics, mode, a, temperature, eps, p, y, out, arg, grad_out, B, T, grad_y, stream, g_err,
                                     sizeof(g_err));
}

extern "C" int pystl_b200_plan(const pystl_formula_t* f, int64_t T, int32_t D, int32_t backward, pystl_plan_t* plan) {
  if (!plan) return set_err(PYSTL_E_INVALID, "plan is NULL");
  Lowered lw;
  TileGeom g;
  int rc = prepare(f, 1, T, D, backward != 0, &lw, &g);
  // formulas only the fused tier evaluates (weights, non-classical Until gradients, ...)
  if (rc == PYSTL_E_UNSUPPORTED && tier_mode() != TIER_INTERP) return pystl_b200_fused_plan(f, T, D, backward, plan);
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

extern "C" int pystl_b200_eval_forward(const pystl_formula_t* f, const float* signal, int64_t B, int64_t T, int32_t D,
                                       float* out, double* mean_acc, void* stream) {
  Lowered lw;
  TileGeom g;
  int rc;
  const int tier = tier_mode();
  if (tier != TIER_INTERP) {
    if (B < 0) return set_err(PYSTL_E_INVALID, "bad shape");
    std::shared_ptr<Prepared> p = fused_prepared(f, T, D, false);
    rc = p->rc;
    if (rc == PYSTL_OK) {
      if (B == 0 || T == 0) return PYSTL_OK;
      if (!signal || !out) return set_err(PYSTL_E_INVALID, "signal / out is NULL");
      fused_bind(*p, B, signal, nullptr, &g);
      void* args[] = {(void*)&p->lw.prog, (void*)&g, (void*)&signal, (void*)&out, (void*)&mean_acc};
      rc = fused_launch(*p, g, args, reinterpret_cast<cudaStream_t>(stream));
    } else {
      snprintf(g_err, sizeof(g_err), "%s", p->err.c_str());
    }
    if (rc != PYSTL_E_UNSUPPORTED || tier == TIER_FUSED) return rc;
  }
  rc = prepare(f, B, T, D, false, &lw, &g);
  if (rc) return rc;
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!signal || !out) return set_err(PYSTL_E_INVALID, "signal / out is NULL");
  if (g.nt % 32 != 0) return set_err(PYSTL_E_INVALID, "CTA size %d is not a multiple of 32", g.nt);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (lw.prog.sem) {
    case SEM_CLASSIC: return launch_fwd_sem<SEM_CLASSIC>(lw, g, signal, out, mean_acc, s, g_err, (int)sizeof(g_err));
    case SEM_SMOOTH: return launch_fwd_sem<SEM_SMOOTH>(lw, g, signal, out, mean_acc, s, g_err, (int)sizeof(g_err));
    case SEM_CUMULATIVE: return launch_fwd_sem<SEM_CUMULATIVE>(lw, g, signal, out, mean_acc, s, g_err, (int)sizeof(g_err));
    case SEM_AGM: return launch_fwd_sem<SEM_AGM>(lw, g, signal, out, mean_acc, s, g_err, (int)sizeof(g_err));
    case SEM_DGMSR: return launch_fwd_sem<SEM_DGMSR>(lw, g, signal, out, mean_acc, s, g_err, (int)sizeof(g_err));
    default: return set_err(PYSTL_E_INVALID, "unknown semantics");
  }
}

// value (optional) and gradient through ONE fused kernel; PYSTL_E_UNSUPPORTED: not on the fused tier
static int fused_value_and_grad(const pystl_formula_t* f, const float* signal, int64_t B, int64_t T, int32_t D,
                                const float* grad_out, float grad_scale, float* grad_signal, float* out, double* mean_acc,
                                cudaStream_t stream) {
  if (B < 0) return set_err(PYSTL_E_INVALID, "bad shape");
  std::shared_ptr<Prepared> p = fused_prepared(f, T, D, true);
  if (p->rc) {
    snprintf(g_err, sizeof(g_err), "%s", p->err.c_str());
    return p->rc;
  }
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!signal || !grad_signal) return set_err(PYSTL_E_INVALID, "signal / grad_signal is NULL");
  TileGeom g;
  fused_bind(*p, B, signal, grad_signal, &g);
  void* args[] = {(void*)&p->lw.prog, (void*)&g,           (void*)&signal, (void*)&grad_out,
                  (void*)&grad_scale, (void*)&grad_signal, (void*)&out,    (void*)&mean_acc};
  return fused_launch(*p, g, args, stream);
}

extern "C" int pystl_b200_eval_backward(const pystl_formula_t* f, const float* signal, int64_t B, int64_t T, int32_t D,
                                        const float* grad_out, float grad_scale, float* grad_signal, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int tier = tier_mode();
  int rc;
  if (tier != TIER_INTERP) {
    rc = fused_value_and_grad(f, signal, B, T, D, grad_out, grad_scale, grad_signal, nullptr, nullptr, s);
    if (rc != PYSTL_E_UNSUPPORTED || tier == TIER_FUSED) return rc;
  }
  Lowered lw;
  TileGeom g;
  rc = prepare(f, B, T, D, true, &lw, &g);
  if (rc) return rc;
  if (B == 0 || T == 0) return PYSTL_OK;
  if (!signal || !grad_signal) return set_err(PYSTL_E_INVALID, "signal / grad_signal is NULL");
  if (g.nt % 32 != 0) return set_err(PYSTL_E_INVALID, "CTA size %d is not a multiple of 32", g.nt);
  const int el = (int)sizeof(g_err);
  switch (lw.prog.sem) {
    case SEM_CLASSIC: return launch_bwd_sem<SEM_CLASSIC>(lw, g, signal, grad_out, grad_scale, grad_signal, s, g_err, el);
    case SEM_SMOOTH: return launch_bwd_sem<SEM_SMOOTH>(lw, g, signal, grad_out, grad_scale, grad_signal, s, g_err, el);
    case SEM_AGM: return launch_bwd_sem<SEM_AGM>(lw, g, signal, grad_out, grad_scale, grad_signal, s, g_err, el);
    case SEM_DGMSR: return launch_bwd_sem<SEM_DGMSR>(lw, g, signal, grad_out, grad_scale, grad_signal, s, g_err, el);
    default: return set_err(PYSTL_E_UNSUPPORTED, "no backward kernel for this semantics");
  }
}

// Robustness AND gradient in one pass over the signal (what Formula.evaluate_with_grad returns,
// api.py:64-78, for every t at once): out / mean_acc as in eval_forward, grad_signal as in
// eval_backward.  One fused kernel when the formula is on the fused tier (the signal is read
// once, nothing is recomputed), otherwise the forward and the backward kernel back to back.
extern "C" int pystl_b200_eval_value_and_grad(const pystl_formula_t* f, const float* signal, int64_t B, int64_t T, int32_t D,
                                              const float* grad_out, float grad_scale, float* out, double* mean_acc,
                                              float* grad_signal, void* stream) {
  if (!out && B > 0 && T > 0) return set_err(PYSTL_E_INVALID, "out is NULL");
  const int tier = tier_mode();
  if (tier != TIER_INTERP) {
    int rc = fused_value_and_grad(f, signal, B, T, D, grad_out, grad_scale, grad_signal, out, mean_acc,
                                  reinterpret_cast<cudaStream_t>(stream));
    if (rc != PYSTL_E_UNSUPPORTED || tier == TIER_FUSED) return rc;
  }
  int rc = pystl_b200_eval_forward(f, signal, B, T, D, out, mean_acc, stream);
  if (rc) return rc;
  return pystl_b200_eval_backward(f, signal, B, T, D, grad_out, grad_scale, grad_signal, stream);
}
