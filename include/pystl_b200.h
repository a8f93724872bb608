/*
 * pystl_b200 -- C ABI of the B200-native robustness evaluator.
 *
 * This is the drop-in boundary for PySTL's robustness hot path.  The reference
 * (ericpalto/PySTL 1.0.0) is pure Python and has no native interface, so every
 * entry point below names the Python interface it stands in for:
 *
 *   reference interface                                   | entry point here
 *   ------------------------------------------------------+----------------------------------
 *   Formula.evaluate(signal, semantics, t)                | pystl_b200_eval_forward
 *     pystl/api.py:142-302 (recursive driver) with the    |   (all t of all traces in one call)
 *     operator interface pystl/semantics/base.py:18-75    |
 *   backward of the torch backend (autograd through       | pystl_b200_eval_backward
 *     pystl/torch/semantics.py:87-587), and               |
 *     DgmsrSemantics.evaluate_with_grad,                  |
 *     pystl/semantics/dgmsr.py:120-274                    |
 *   Interval / _window_indices clipping, api.py:29-54;    | pystl_b200_plan (validity, tiling)
 *     "window is empty" ValueError, api.py:218-221        |
 *   create_semantics(**kwargs): temperature / eps / p,    | pystl_formula_t.{temperature,eps,p}
 *     pystl/semantics/__init__.py:108-125                 |
 *
 * Conventions: plain pointers and sizes only; the CALLER owns every buffer
 * (kernels never allocate); device pointers unless the name says host; all
 * launches are asynchronous on `stream` (a cudaStream_t passed as void*);
 * functions return 0 on success or a negative pystl_status; the message for
 * the last failure on the calling thread is pystl_b200_last_error().
 *
 * The binding a maintainer adds on the reference side (ctypes) is shown in
 * INTEGRATION.md; pystl_b200/capi.py is that binding.
 */
#ifndef PYSTL_B200_H
#define PYSTL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYSTL_B200_ABI_VERSION 2
#define PYSTL_MAX_TERMS 8 /* affine terms per predicate */

typedef enum pystl_status {
  PYSTL_OK = 0,
  PYSTL_E_INVALID = -1,     /* malformed formula / arguments (Python: ValueError)            */
  PYSTL_E_UNSUPPORTED = -2, /* valid but outside this build's limits (NotImplementedError)  */
  PYSTL_E_CUDA = -3         /* a CUDA call failed                                            */
} pystl_status;

/* registry syntaxes, pystl/semantics/__init__.py:47-52 */
typedef enum pystl_semantics {
  PYSTL_SEM_CLASSIC = 0,    /* semantics/classic.py    */
  PYSTL_SEM_SMOOTH = 1,     /* semantics/smooth.py     */
  PYSTL_SEM_CUMULATIVE = 2, /* semantics/cumulative.py */
  PYSTL_SEM_AGM = 3,        /* semantics/agm.py        */
  PYSTL_SEM_DGMSR = 4       /* semantics/dgmsr.py, _dgmsr.py */
} pystl_semantics;

/* node classes of pystl/api.py:129-302 */
typedef enum pystl_node_kind {
  PYSTL_NODE_PRED = 0,
  PYSTL_NODE_NOT = 1,
  PYSTL_NODE_AND = 2,
  PYSTL_NODE_OR = 3,
  PYSTL_NODE_ALWAYS = 4,
  PYSTL_NODE_EVENTUALLY = 5,
  PYSTL_NODE_UNTIL = 6
} pystl_node_kind;

/* One formula node.  Nodes are listed in post-order (children before parents);
 * the root is the last node. */
typedef struct pystl_node {
  int32_t kind;       /* pystl_node_kind */
  int32_t n_children; /* PRED 0, NOT 1, AND/OR n, ALWAYS/EVENTUALLY 1, UNTIL 2 (left, right) */
  int32_t child_off;  /* first entry of this node's children in pystl_formula_t.children */
  int32_t a;          /* Interval.start (api.py:33) */
  int32_t b;          /* Interval.end, -1 = None = up to the horizon (api.py:34,51) */
  int32_t pred;       /* PRED: index into pystl_formula_t.preds */
  int32_t w_off;      /* weights: AND/OR per child; ALWAYS/EVENTUALLY per window position;
                         UNTIL weights_left.  -1 = None (uniform) */
  int32_t w_len;
  int32_t w2_off;     /* UNTIL weights_right, -1 = None */
  int32_t w2_len;
  float wp0, wp1;     /* UNTIL weights_pair (api.py:265), default (1, 1) */
} pystl_node_t;

/* Affine predicate rho(t) = bias + sum_k coeff[k] * signal[t, channel[k]]
 * (the structured form of Predicate.fn, api.py:133-143). */
typedef struct pystl_pred {
  int32_t n_terms;
  int32_t channel[PYSTL_MAX_TERMS];
  float coeff[PYSTL_MAX_TERMS];
  float bias;
  int32_t slack; /* ABI 2.  0 for a predicate on the signal.  > 0: the channel holds values the host computed
                    (pystl_b200_open_window_forward) that do not exist at t > T-1-slack -- the reference raises
                    there (api.py:218-221) and the outputs that depend on them are NaN; -1: it raises for every t */
  int32_t pair;  /* ABI 2, cumulative only.  1: the value is the PAIR (pos, neg) = (signal[t, channel[0]], signal[t, channel[1]])
                    as it stands (n_terms = 2; coeff and bias unused) instead of (max(0, rho), min(0, rho)) of an affine rho
                    (cumulative.py:24-30): how a host-computed cumulative value (an open-ended window) enters a formula */
} pystl_pred_t;

typedef struct pystl_formula {
  int32_t semantics;  /* pystl_semantics */
  float temperature;  /* smooth: SmoothRobustSemantics(temperature=), smooth.py:28 */
  float eps;          /* dgmsr: DgmsrSemantics(eps=), dgmsr.py:31 */
  int32_t p;          /* dgmsr: DgmsrSemantics(p=),   dgmsr.py:31 */
  int32_t n_nodes;
  const pystl_node_t* nodes;
  int32_t n_children;
  const int32_t* children; /* node indices */
  int32_t n_preds;
  const pystl_pred_t* preds;
  int32_t n_weights;
  const float* weights;
} pystl_formula_t;

/* What pystl_b200_plan reports (host only, no GPU needed). */
typedef struct pystl_plan {
  int32_t out_channels;  /* 1; 2 for cumulative (pos, neg), cumulative.py:13-18 */
  int32_t lookahead;     /* samples after t the value at t depends on; -1 = unbounded */
  int32_t start_slack;   /* outputs at t > T-1-start_slack have an empty window: the
                            reference raises ValueError (api.py:218-221); written as NaN */
  int32_t raises_always; /* 1: an UNTIL child has start_slack > 0, the reference raises for
                            every t (api.py:288-293) */
  int32_t threads;       /* CTA size chosen for this (formula, T, D) */
  int32_t tile_ext;      /* samples staged per CTA (= 8 * threads) */
  int32_t tile_out;      /* outputs produced per CTA */
  int32_t n_tiles;       /* tiles per trace */
  int32_t smem_bytes;    /* dynamic shared memory per CTA */
  int32_t n_instr;       /* lowered tile instructions */
  int32_t n_planes;      /* shared-memory value planes */
} pystl_plan_t;

int pystl_b200_abi_version(void);
const char* pystl_b200_last_error(void);

/* Validate + lower the formula for traces of length T with D channels.
 * backward != 0 plans the gradient kernel instead of the forward one. */
int pystl_b200_plan(const pystl_formula_t* formula, int64_t T, int32_t D, int32_t backward,
                    pystl_plan_t* plan);

/* Robustness of every (trace, t):  out[b, t, c] for signal[b, t, d] (fp32, contiguous).
 *   out      [B, T, out_channels]  written for all t; NaN where the reference raises
 *   mean_acc optional double[2]: += (sum of the valid outputs' channels, count); the
 *            fused mean-robustness reduction (one NCCL all-reduce of these 16 bytes
 *            finishes a batch-sharded mean).  May be NULL.
 * Replaces the per-(trace, t) recursion of pystl/api.py:142-302. */
int pystl_b200_eval_forward(const pystl_formula_t* formula, const float* signal, int64_t B,
                            int64_t T, int32_t D, float* out, double* mean_acc, void* stream);

/* Gradient of sum_{b,t,c} grad_out[b,t,c] * out[b,t,c] with respect to signal.
 *   grad_out    [B, T, out_channels] or NULL for the constant cotangent grad_scale
 *               (d mean / d out); entries at t the reference raises on are ignored
 *   grad_signal [B, T, D], fully overwritten (zeros for unreferenced channels)
 * Replaces autograd through pystl/torch/semantics.py and
 * DgmsrSemantics.evaluate_with_grad (pystl/semantics/dgmsr.py:120-274). */
int pystl_b200_eval_backward(const pystl_formula_t* formula, const float* signal, int64_t B,
                             int64_t T, int32_t D, const float* grad_out, float grad_scale,
                             float* grad_signal, void* stream);

/* Robustness AND gradient in one pass over the signal: what Formula.evaluate_with_grad
 * (pystl/api.py:64-78 -> DgmsrSemantics.evaluate_with_grad, pystl/semantics/dgmsr.py:120-274)
 * returns, for every t of every trace at once.  out / mean_acc as in eval_forward,
 * grad_out / grad_scale / grad_signal as in eval_backward.  Formulas on the fused tier run
 * ONE kernel (the signal is read once, no value is recomputed); the others run the forward
 * and the backward kernel back to back. */
int pystl_b200_eval_value_and_grad(const pystl_formula_t* formula, const float* signal, int64_t B,
                                   int64_t T, int32_t D, const float* grad_out, float grad_scale,
                                   float* out, double* mean_acc, float* grad_signal, void* stream);

/* Fused tier (run-time specialised kernels, pystl_b200/csrc/codegen.cpp + jit.cpp).
 * precompile: generate + compile the kernel of (formula, T, D) into the on-disk cubin cache;
 * needs no GPU.  PYSTL_E_UNSUPPORTED: this formula runs on the interpreter tier.
 * source: the generated CUDA source (inspection); returns its length. */
int pystl_b200_fused_precompile(const pystl_formula_t* formula, int64_t T, int32_t D, int32_t backward);
/* tile geometry of the fused kernel (threads, tile_ext, tile_out, n_tiles, smem_bytes) */
int pystl_b200_fused_plan(const pystl_formula_t* formula, int64_t T, int32_t D, int32_t backward,
                          pystl_plan_t* plan);
int pystl_b200_fused_source(const pystl_formula_t* formula, int64_t T, int32_t D, int32_t backward,
                            char* buf, int32_t len);

/* Open-ended windows on traces longer than one tile (ABI 2).  G[a, inf) / F[a, inf) of a VALUE TRACE y[B, T]
 * (the operand's robustness, produced by eval_forward): out[b, t] = the semantics' temporal_always (mode 0) /
 * temporal_eventually (mode 1) of y[b, t+a .. T-1] -- Always / Eventually.evaluate with Interval.end None,
 * pystl/api.py:49-54,208-243 -> classic.py:39-57, smooth.py:47-65, agm.py:92-110, dgmsr.py:58-86 -- as one suffix
 * scan per trace (pystl_b200/csrc/scan.cu).  The host feeds `out` back to the formula kernels as a leaf channel
 * with pystl_pred_t.slack = a.  Outputs at t > T-1-a (empty window) hold a finite filler for that kernel to read; mode + 2
 * writes NaN there instead (the window is the whole formula: NaN is what eval_forward writes where the reference raises).
 *   arg   int32[B, T], classical only (may be NULL otherwise): the window's arg-min / arg-max for the gradient
 * PYSTL_SEM_CUMULATIVE works on ONE channel (pos or neg) of the pairs per call: Always is that channel's minimum
 * (cumulative.py:35-46; needs arg), Eventually its SUM over the window (cumulative.py:62-73); the parent formula reads the two
 * result channels through a pystl_pred_t with pair = 1.  Window weights are not offered here. */
int pystl_b200_open_window_forward(int32_t semantics, int32_t mode, int32_t a, float temperature, float eps, int32_t p,
                                   const float* y, int64_t B, int64_t T, float* out, int32_t* arg, void* stream);
/* grad_y[B, T] (fully overwritten) = d sum(grad_out * out) / d y over the outputs that exist (t <= T-1-a). */
int pystl_b200_open_window_backward(int32_t semantics, int32_t mode, int32_t a, float temperature, float eps, int32_t p,
                                    const float* y, const float* out, const int32_t* arg, const float* grad_out,
                                    int64_t B, int64_t T, float* grad_y, void* stream);

/* Classical Until without an interval end on traces longer than one tile (ABI 2): out[b, t] = W(t + a),
 *   W(u) = max_{t' >= u} min( r[t'], min_{u <= t'' <= t'} l[t''] )            (classic.py:59-80 with Interval.end None)
 * for the value traces l[B, T], r[B, T] of the two operands -- a suffix scan of clamp compositions.  The host wraps it as
 * U[a, inf) = min( G[0, a] l, W(t + a) ) (the decomposition of pystl_b200/csrc/lower.cpp).  key[B, T] records where each
 * value came from (2 * sample + 1 for an r, + 0 for an l); backward routes the cotangents there (grad_l, grad_r fully
 * overwritten).  The other semantics' Until is O(T^2) per trace without an end and stays inside one tile
 * (PYSTL_E_UNSUPPORTED here). */
int pystl_b200_open_until_forward(int32_t semantics, int32_t a, const float* l, const float* r, int64_t B, int64_t T,
                                  float* out, int32_t* key, void* stream);
int pystl_b200_open_until_backward(int32_t semantics, int32_t a, const int32_t* key, const float* grad_out, int64_t B,
                                   int64_t T, float* grad_l, float* grad_r, void* stream);

/* The tensor a formula kernel reads when some predicates are leaf channels (host-computed columns: opaque predicates,
 * pystl_b200/compiler.py: leaf_values; the open-ended windows above): for every row r = (trace, t)
 *   out[r, :] = [ signal[r, channels[0 .. n_channels)] | cols[0][r] .. cols[n_cols - 1][r] ]       (one pass)
 * and its adjoint: grad_signal[rows, D] fully overwritten (zeros for channels not read; may be NULL), grad_cols[j][rows]
 * (entries may be NULL).  At most 32 read channels of a signal of at most 64, and 16 columns (PYSTL_E_UNSUPPORTED beyond). */
int pystl_b200_assemble_channels(const float* signal, int64_t rows, int32_t D, const int32_t* channels, int32_t n_channels,
                                 const float* const* cols, int32_t n_cols, float* out, void* stream);
int pystl_b200_scatter_channels(const float* grad, int64_t rows, int32_t D, const int32_t* channels, int32_t n_channels,
                                float* grad_signal, float* const* grad_cols, int32_t n_cols, void* stream);

/* CT-STL on value traces (CtstlSemantics, /root/reference/pystl/semantics/ctstl.py -- exported by the reference but not
 * registered under create_semantics, semantics/__init__.py:47-74).  Everything else of CT-STL is the classical semantics.
 *
 * The cumulative-time operator C^tau_I (CtstlSemantics.temporal_cumulative, ctstl.py:107-127, with kth_largest :18-26 and
 * tau_to_k :28-35): out[b, t] = k-th largest of y[b, t + a .. min(t + b_end, T - 1)] (window clipped as api.py:49-54; b_end < 0:
 * no end), k = ceil(tau / delta) >= 1 computed by the caller; -inf where the clipped window holds fewer than k samples.
 * arg[B, T] (may be NULL) records the sample the value came from (-1: none); among equal samples the one a stable ascending
 * sort puts at position n - k (jax/semantics.py:378-386; -0 and +0 are equal samples, a selected -0 is returned as +0).  A NaN
 * inside a window (an operand value the reference raises on) makes that output NaN.  Windows of at most 49152 samples
 * (PYSTL_E_UNSUPPORTED beyond).
 * backward: grad_y[B, T] (fully overwritten) = d sum(grad_out * out) / d y. */
int pystl_b200_ctstl_cumulative_forward(const float* y, int64_t B, int64_t T, int32_t a, int32_t b_end, int32_t k, float* out,
                                        int32_t* arg, void* stream);
int pystl_b200_ctstl_cumulative_backward(const int32_t* arg, const float* grad_out, int64_t B, int64_t T, int32_t a, int32_t b_end,
                                         float* grad_y, void* stream);
/* CT-STL Until (CtstlSemantics.temporal_until, ctstl.py:78-105): the left prefix of candidate i EXCLUDES t + i,
 *   out[b, t] = max_{i = a .. min(b_end, T - 1 - t)} min( r[t + i], min_{k < i} l[t + k] )       (b_end < 0: no end)
 * on the value traces l[B, T], r[B, T]; NaN where the window is empty (the reference raises ValueError, ctstl.py:96-97).
 * key[B, T] (may be NULL): 2 * sample + 1 when the value is an r, 2 * sample when it is an l, -1 where there is none;
 * backward routes the cotangents there (grad_l, grad_r fully overwritten). */
int pystl_b200_ctstl_until_forward(const float* l, const float* r, int64_t B, int64_t T, int32_t a, int32_t b_end, float* out,
                                   int32_t* key, void* stream);
int pystl_b200_ctstl_until_backward(const int32_t* key, const float* grad_out, int64_t B, int64_t T, int32_t a, int32_t b_end,
                                    float* grad_l, float* grad_r, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PYSTL_B200_H */
