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
