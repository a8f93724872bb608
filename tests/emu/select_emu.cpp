// CPU build of the CT-STL kernels (pystl_b200/csrc/select.cu) for the test tier without a GPU (TEST INFRASTRUCTURE): the very
// same kernel source, every CUDA thread a fiber of simt_host.h.  Grids are walked CTA by CTA; the grid-stride kernels run on a
// deliberately small grid so that the stride loop is exercised.
#include "simt_host.h"
#define PYSTL_SELECT_EMU 1
#include "../../pystl_b200/csrc/select.cu"

namespace {
using namespace pystl;
struct KthArgs { const float* y; float* out; int* arg; int T, a, W, k, tiles; };
void kth_body(void* p) { auto* A = static_cast<KthArgs*>(p); kth_window_fwd(A->y, A->out, A->arg, A->T, A->a, A->W, A->k, A->tiles); }
struct KthBwdArgs { const int* arg; const float* go; float* gy; long long B; int T, a, W; };
void kth_bwd_body(void* p) { auto* A = static_cast<KthBwdArgs*>(p); kth_window_bwd(A->arg, A->go, A->gy, A->B, A->T, A->a, A->W); }
struct UntilArgs { const float* l; const float* r; float* out; int* key; long long B; int T, a, b; };
void until_body(void* p) { auto* A = static_cast<UntilArgs*>(p); ctstl_until_fwd(A->l, A->r, A->out, A->key, A->B, A->T, A->a, A->b); }
struct UntilBwdArgs { const int* key; const float* go; float* gl; float* gr; long long B; int T, a, b; };
void until_bwd_body(void* p) { auto* A = static_cast<UntilBwdArgs*>(p); ctstl_until_bwd(A->key, A->go, A->gl, A->gr, A->B, A->T, A->a, A->b); }
long long window_len(int a, int b_end, long long T) {
  long long W = b_end >= 0 ? (long long)b_end - a + 1 : T - a;
  if (W > T) W = T;
  return W < 1 ? 1 : W;
}
void run_grid(unsigned grid, int nt, size_t smem, void (*body)(void*), void* arg) {
  sel_emu::grid_dim() = grid;
  for (unsigned b = 0; b < grid; ++b) simt::run_cta((int)b, 0, nt, smem, body, arg);
}
}  // namespace

extern "C" int pystl_select_emu_cumulative_forward(const float* y, long long B, long long T, int a, int b_end, int k, float* out, int* arg) {
  if (B == 0 || T == 0) return 0;
  const long long W = window_len(a, b_end, T);
  const long long tiles = (T + kSelOut - 1) / kSelOut;
  KthArgs A{y, out, arg, (int)T, a, (int)W, k, (int)tiles};
  run_grid((unsigned)(B * tiles), kSelThreads, (size_t)((kSelOut + W + 6) & ~3ll) * 4, kth_body, &A);
  return 0;
}
extern "C" int pystl_select_emu_cumulative_backward(const int* arg, const float* go, long long B, long long T, int a, int b_end, float* gy) {
  if (B == 0 || T == 0) return 0;
  KthBwdArgs A{arg, go, gy, B, (int)T, a, (int)window_len(a, b_end, T)};
  run_grid(3, 256, 16, kth_bwd_body, &A);
  return 0;
}
extern "C" int pystl_select_emu_until_forward(const float* l, const float* r, long long B, long long T, int a, int b_end, float* out, int* key) {
  if (B == 0 || T == 0) return 0;
  UntilArgs A{l, r, out, key, B, (int)T, a, b_end};
  run_grid(3, 256, 16, until_body, &A);
  return 0;
}
extern "C" int pystl_select_emu_until_backward(const int* key, const float* go, long long B, long long T, int a, int b_end, float* gl, float* gr) {
  if (B == 0 || T == 0) return 0;
  UntilBwdArgs A{key, go, gl, gr, B, (int)T, a, b_end};
  run_grid(3, 256, 16, until_bwd_body, &A);
  return 0;
}
