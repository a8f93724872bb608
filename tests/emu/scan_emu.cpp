// CPU build of the open-ended scan kernels (pystl_b200/csrc/scan.cu) for the test tier without a GPU (TEST INFRASTRUCTURE): the
// very same kernel source, every CUDA thread a fiber of simt_host.h, one warp per trace exactly as on the device.  The entry points
// mirror pystl_b200_open_window_forward / _backward and pystl_b200_open_until_forward / _backward (include/pystl_b200.h) minus the
// stream; the dispatch over (semantics, mode) follows scan.cu's launchers.
#include "simt_host.h"
#define PYSTL_SCAN_EMU 1
#include "../../pystl_b200/csrc/scan.cu"

namespace {
using namespace pystl;
struct FwdArgs { const float* y; float* out; int* arg; long long B; int T, a; RP rp; int nan_tail; };
template <int SEM, int MODE>
void fwd_body(void* p) { auto* A = static_cast<FwdArgs*>(p); open_window_fwd<SEM, MODE>(A->y, A->out, A->arg, A->B, A->T, A->a, A->rp, A->nan_tail); }
struct BwdArgs { const float* y; const float* out; const float* go; float* gy; long long B; int T, a; RP rp; };
template <int SEM, int MODE>
void bwd_body(void* p) { auto* A = static_cast<BwdArgs*>(p); open_window_bwd<SEM, MODE>(A->y, A->out, A->go, A->gy, A->B, A->T, A->a, A->rp); }
struct SegArgs { const int* arg; const float* go; float* gy; float* gr; long long B; int T, a; };
template <bool TWO>
void seg_body(void* p) { auto* A = static_cast<SegArgs*>(p); open_window_bwd_classic<TWO>(A->arg, A->go, A->gy, A->gr, A->B, A->T, A->a); }
struct UntilArgs { const float* l; const float* r; float* out; int* key; long long B; int T, a; };
void until_body(void* p) { auto* A = static_cast<UntilArgs*>(p); open_until_fwd_classic(A->l, A->r, A->out, A->key, A->B, A->T, A->a); }

void run_grid(long long B, void (*body)(void*), void* arg) {
  const unsigned grid = (unsigned)((B + kWarpsPerBlock - 1) / kWarpsPerBlock);
  for (unsigned b = 0; b < grid; ++b) simt::run_cta((int)b, 0, 32 * kWarpsPerBlock, 16, body, arg);
}
#define SCAN_EMU_PICK(BODY, sem, mode)                                                         \
  ((sem) == SEM_CLASSIC ? ((mode) == MODE_AND ? BODY<SEM_CLASSIC, MODE_AND> : BODY<SEM_CLASSIC, MODE_OR>) \
   : (sem) == SEM_SMOOTH ? ((mode) == MODE_AND ? BODY<SEM_SMOOTH, MODE_AND> : BODY<SEM_SMOOTH, MODE_OR>) \
   : (sem) == SEM_AGM    ? ((mode) == MODE_AND ? BODY<SEM_AGM, MODE_AND> : BODY<SEM_AGM, MODE_OR>)       \
                         : ((mode) == MODE_AND ? BODY<SEM_DGMSR, MODE_AND> : BODY<SEM_DGMSR, MODE_OR>))
// (the adjoint of the classical window is the segmented-run kernel: open_window_bwd has no classical instance)
#define SCAN_EMU_PICK_NC(BODY, sem, mode)                                                      \
  ((sem) == SEM_SMOOTH ? ((mode) == MODE_AND ? BODY<SEM_SMOOTH, MODE_AND> : BODY<SEM_SMOOTH, MODE_OR>) \
   : (sem) == SEM_AGM  ? ((mode) == MODE_AND ? BODY<SEM_AGM, MODE_AND> : BODY<SEM_AGM, MODE_OR>)       \
                       : ((mode) == MODE_AND ? BODY<SEM_DGMSR, MODE_AND> : BODY<SEM_DGMSR, MODE_OR>))
}  // namespace

extern "C" int pystl_scan_emu_window_forward(int semantics, int mode, int a, float temperature, float eps, int p, const float* y,
                                             long long B, long long T, float* out, int* arg) {
  if (B == 0 || T == 0) return 0;
  if (semantics == SEM_CUMULATIVE && (mode & 1) == 0) semantics = SEM_CLASSIC;  // (scan.cu: open_window_forward)
  FwdArgs A{y, out, semantics == SEM_CLASSIC ? arg : nullptr, B, (int)T, a, make_rp(temperature, eps, p), (mode >> 1) & 1};
  mode &= 1;
  if (semantics == SEM_CUMULATIVE) run_grid(B, fwd_body<SEM_CUMULATIVE, MODE_OR>, &A);
  else run_grid(B, SCAN_EMU_PICK(fwd_body, semantics, mode), &A);
  return 0;
}
extern "C" int pystl_scan_emu_window_backward(int semantics, int mode, int a, float temperature, float eps, int p, const float* y,
                                              const float* out, const int* arg, const float* go, long long B, long long T, float* gy) {
  if (B == 0 || T == 0) return 0;
  mode &= 1;
  if (semantics == SEM_CUMULATIVE && mode == 0) semantics = SEM_CLASSIC;
  if (semantics == SEM_CLASSIC) {
    memset(gy, 0, (size_t)B * (size_t)T * sizeof(float));  // (cudaMemsetAsync in the product)
    SegArgs S{arg, go, gy, nullptr, B, (int)T, a};
    run_grid(B, seg_body<false>, &S);
    return 0;
  }
  BwdArgs A{y, out, go, gy, B, (int)T, a, make_rp(temperature, eps, p)};
  if (semantics == SEM_CUMULATIVE) run_grid(B, bwd_body<SEM_CUMULATIVE, MODE_OR>, &A);
  else run_grid(B, SCAN_EMU_PICK_NC(bwd_body, semantics, mode), &A);
  return 0;
}
extern "C" int pystl_scan_emu_until_forward(int a, const float* l, const float* r, long long B, long long T, float* out, int* key) {
  if (B == 0 || T == 0) return 0;
  UntilArgs A{l, r, out, key, B, (int)T, a};
  run_grid(B, until_body, &A);
  return 0;
}
extern "C" int pystl_scan_emu_until_backward(int a, const int* key, const float* go, long long B, long long T, float* gl, float* gr) {
  if (B == 0 || T == 0) return 0;
  memset(gl, 0, (size_t)B * (size_t)T * sizeof(float));
  memset(gr, 0, (size_t)B * (size_t)T * sizeof(float));
  SegArgs S{key, go, gl, gr, B, (int)T, a};
  run_grid(B, seg_body<true>, &S);
  return 0;
}

// ---- the tensor a formula kernel reads next to leaf channels (scan.cu: assemble_rows / scatter_rows), on a deliberately small grid
namespace {
struct AsmArgs { const float* signal; AssembleArgs a; float* out; long long rows; };
template <int WT>
void asm_body(void* p) { auto* A = static_cast<AsmArgs*>(p); assemble_rows<WT>(A->signal, A->a, A->out, A->rows); }
struct SctArgs { const float* grad; AssembleArgs a; float* gs; long long rows; };
template <int DT>
void sct_body(void* p) { auto* A = static_cast<SctArgs*>(p); scatter_rows<DT>(A->grad, A->a, A->gs, A->rows); }
void fill(AssembleArgs& a, int D, const int* channels, int n, int L) {  // (scan.cu: assemble_fill, minus the validation)
  a.n = n, a.L = L, a.D = D;
  for (int d = 0; d < kMaxAssembleD; ++d) a.inv[d] = -1;
  for (int i = 0; i < n; ++i) a.ch[i] = channels[i], a.inv[channels[i]] = (signed char)i;
}
void run_rows(unsigned grid, void (*body)(void*), void* arg) {
  scan_emu::grid_dim() = grid;
  for (unsigned b = 0; b < grid; ++b) simt::run_cta((int)b, 0, 256, 16, body, arg);
}
}  // namespace

extern "C" int pystl_scan_emu_assemble(const float* signal, long long rows, int D, const int* channels, int n, const float* const* cols,
                                       int L, float* out, int templated) {
  AsmArgs A{signal, AssembleArgs{}, out, rows};
  fill(A.a, D, channels, n, L);
  for (int j = 0; j < L; ++j) A.a.cols[j] = cols[j];
  void (*body)(void*) = asm_body<0>;
  if (templated) switch (n + L) {
#define EMU_ASM(WT) case WT: body = asm_body<WT>; break;
    EMU_ASM(1) EMU_ASM(2) EMU_ASM(3) EMU_ASM(4) EMU_ASM(5) EMU_ASM(6) EMU_ASM(7) EMU_ASM(8)
#undef EMU_ASM
    default: break;
  }
  run_rows(3, body, &A);
  return 0;
}
extern "C" int pystl_scan_emu_scatter(const float* grad, long long rows, int D, const int* channels, int n, float* grad_signal,
                                      float* const* gcols, int L, int templated) {
  SctArgs A{grad, AssembleArgs{}, grad_signal, rows};
  fill(A.a, D, channels, n, L);
  for (int j = 0; j < L; ++j) A.a.gcols[j] = gcols[j];
  void (*body)(void*) = sct_body<0>;
  if (templated) switch (D) {
#define EMU_SCT(DT) case DT: body = sct_body<DT>; break;
    EMU_SCT(1) EMU_SCT(2) EMU_SCT(3) EMU_SCT(4) EMU_SCT(5) EMU_SCT(6) EMU_SCT(7) EMU_SCT(8)
#undef EMU_SCT
    default: break;
  }
  run_rows(3, body, &A);
  return 0;
}
