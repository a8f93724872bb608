// Backward pass of the fused formula kernel, single source like tile_eval.h.
//
// A CTA owns the gradient of `lout` consecutive samples of one trace.  It stages the
// signal of those samples plus a halo of H = formula look-ahead samples on BOTH sides
// (outputs up to H samples earlier depend on them; their windows reach H samples
// further), recomputes every node value of the formula in shared memory (no
// intermediate ever went to HBM in the forward pass, so there is nothing to re-read),
// and runs the instruction list backwards on adjoint planes:
//
//   classic   values carry their provenance (SEM_CLASSIC_SRC): the gradient of a min / max
//             tree is one scatter of the root cotangent through the recorded leaf, exactly
//             the routing of torch.min / max(dim) in the reference (torch/semantics.py:106,113)
//   smooth    d softmin / d x_q = exp((out_t - x_q) / tau) = e_q / S_t in the block-shifted
//             forward quantities, so the adjoint of a window is a backward-looking window SUM
//             of g_t / S_t times e_q ("Smooth backward"); windows whose
//             shifted sum underflowed are scattered exactly, one term at a time
//   agm /     branch-masked partials (agm.py:60-88, _dgmsr.py:80-106): two backward-looking
//   dgmsr     window sums (arithmetic-branch and geometric-branch coefficients of every
//             output) times a per-sample factor
//
// The same van Herk machinery as the forward pass does the window sums (offsets
// [-b, -a], plain addition).  Reference for what is differentiated: autograd through
// pystl/torch/semantics.py and the analytic chain rule of pystl/semantics/_dgmsr.py:66-119.
#pragma once
#include "tile_eval.h"

namespace pystl {

template <int AW>
struct SumOp {
  using A = Vec<AW>;
  static PYSTL_HD A ident() {
    A r;
#pragma unroll
    for (int w = 0; w < AW; ++w) r.x[w] = 0.f;
    return r;
  }
  static PYSTL_HD A comb(A a, A b) {
    A r;
#pragma unroll
    for (int w = 0; w < AW; ++w) r.x[w] = a.x[w] + b.x[w];
    return r;
  }
};

// adjoint plane of value slot `slot` (one word per slot; the host assigned it, reusing planes
// whose contents are dead by the time the adjoint is first written: lower.cpp)
template <int SEM>
PYSTL_HD float* adj_plane(const Ctx& c, int slot) {
  return c.smem + c.off_planes + c.adj_tab()[slot] * c.lext;
}

// g_plane[own chunk] (+)= sign * g; `first`: this is the first write, the plane holds garbage
template <int SEM>
PYSTL_HD void adj_accum8(const Ctx& c, int slot, bool neg, bool first, int tid, const float (&g)[8]) {
  float* pl = adj_plane<SEM>(c, slot);
  float cur[8];
  if (first) {
#pragma unroll
    for (int i = 0; i < 8; ++i) cur[i] = 0.f;
  } else {
    load8(pl, tid, cur);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) cur[i] += neg ? -g[i] : g[i];
  store8(pl, tid, cur);
}

// derivative of min(v, 0) / max(v, 0) as torch.minimum / maximum give it (0.5 on the tie)
PYSTL_HD float dmin0(float v) { return v < 0.f ? 1.f : (v == 0.f ? 0.5f : 0.f); }
PYSTL_HD float dmax0(float v) { return v > 0.f ? 1.f : (v == 0.f ? 0.5f : 0.f); }

// ------------------------------------------------------------------ per-semantics partials
// Forward: out = fin(sum_j lift(x_j)).  d out / d x_j factorises into a coefficient that only
// depends on the OUTPUT (KW words, `out_coeff`) and a factor that only depends on the INPUT
// (`in_factor`), which is what lets a window adjoint be a window sum.
template <int SEM, int MODE>
struct Adj;

template <int MODE>
struct Adj<SEM_SMOOTH, MODE> {
  static constexpr int KW = 1;
  static constexpr float kMaxExp = 100.f;
  // K_t = g 2^{M -+ k out}; e_q = 2^{+-k x - M}; product = g exp(+-(x - out)/tau) <= g
  static PYSTL_HD float expo_out(float o, const RP& rp) { return rp.shift - (MODE == MODE_AND ? -o : o) * rp.k; }
  static PYSTL_HD Vec<1> out_coeff(float g, float o, float sw, const RP& rp, bool& lost) {
    const float e = expo_out(o, rp);
    lost = !(e <= kMaxExp);  // the forward sum underflowed here (or out is NaN): exact path
    return Vec<1>{{lost ? 0.f : g * fast_ex2(e)}};
  }
  static PYSTL_HD float in_factor(Vec<1> w, float x, float, const RP& rp) {
    return w.x[0] * fast_ex2((MODE == MODE_AND ? -x : x) * rp.k - rp.shift);
  }
  // exact weight of one (output, input) pair
  static PYSTL_HD float pair_weight(float o, float x, const RP& rp) {
    return fast_ex2((MODE == MODE_AND ? (o - x) : (x - o)) * rp.k);
  }
};

template <int MODE>
struct Adj<SEM_AGM, MODE> {
  static constexpr int KW = 2;
  static PYSTL_HD Vec<2> out_coeff(float g, float o, float sw, const RP&, bool& lost) {
    lost = false;
    const float gn = g * fast_rcp(sw);
    if constexpr (MODE == MODE_AND) return (o <= 0.f) ? Vec<2>{{gn, 0.f}} : Vec<2>{{0.f, gn * (o + 1.f)}};
    // (o == 0 is the arithmetic branch whose positive entries all carry weight 0)
    else return (o >= 0.f) ? Vec<2>{{gn, 0.f}} : Vec<2>{{0.f, gn * (1.f - o)}};
  }
  static PYSTL_HD float in_factor(Vec<2> w, float x, float wt, const RP&) {
    // AND: d min(x,0) = 1 (x<0), 1/2 (x==0, torch.minimum); the log1p branch only for x > 0.  OR mirrored
    // (agm.py:86 tests x > 0: x == 0 belongs to the geometric branch there, with the tie half of max(x,0)).
    if constexpr (MODE == MODE_AND) {
      return wt * (x > 0.f ? w.x[1] * fast_rcp(1.f + x) : (x < 0.f ? w.x[0] : 0.5f * w.x[0]));
    } else {
      return wt * (x > 0.f ? w.x[0] : (x < 0.f ? w.x[1] * fast_rcp(1.f - x) : 0.5f * w.x[0] + w.x[1]));
    }
  }
};

template <int MODE>
struct Adj<SEM_DGMSR, MODE> {
  static constexpr int KW = 2;
  // in the u-domain (u = x for AND, -x for OR; h = out for AND, -out for OR), _dgmsr.py:80-106
  static PYSTL_HD Vec<2> out_coeff(float g, float o, float sw, const RP& rp, bool& lost) {
    lost = false;
    const float h = MODE == MODE_AND ? o : -o;
    const float gs = g * fast_rcp(sw);
    // arithmetic branch (h <= 0): mp^(1/2 - p) / sw with sqrt(mp) = sqrt(eps) - h
    const float rmp = rp.sqrt_eps - h;
    const float cn = PYSTL_P(rp) == 1 ? fast_rcp(rmp) : fast_pow(fmaxf(rmp, 1e-30f), 1.f - 2.f * (float)PYSTL_P(rp));
    // geometric branch: sqrt(m0) (1 - (eps / m0)^sw) / sw with sqrt(m0) = h + sqrt(eps)
    const float rm0 = h + rp.sqrt_eps;
    const float ratio = fast_exp(sw * (rp.ln_eps - 2.f * fast_ln(fmaxf(rm0, 1e-30f))));
    const bool neg = h <= 0.f;
    return Vec<2>{{neg ? gs * cn : 0.f, neg ? 0.f : gs * rm0 * (1.f - ratio)}};
  }
  static PYSTL_HD float in_factor(Vec<2> w, float x, float wt, const RP& rp) {
    const float u = MODE == MODE_AND ? x : -x;
    const float au = fabsf(u);
    const float pw = PYSTL_P(rp) == 1 ? au : (au > 0.f ? fast_pow(au, 2.f * (float)PYSTL_P(rp) - 1.f) : 0.f);
    return wt * ((u <= 0.f) ? w.x[0] * pw : w.x[1] * fast_rcp(au));
  }
};

// ------------------------------------------------------------------ n-ary And / Or, backward
template <int SEM, int MODE, class EX>
PYSTL_HD void op_nary_bwd(EX& ex, const Ctx& c, const TInstr& in) {
  using AD = Adj<SEM, MODE>;
  const int nsrc = in.nsrc, dst = in.dst;
  const unsigned negmask = in.negmask, firstmask = in.adj_first;
  const float sw = in.sw;
  PYSTL_FOR_THREADS(ex, tid) {
    RP rp = c.rp;
    float g[8];
    Vec<1> o[8];
    load8(adj_plane<SEM>(c, dst), tid, g);
    load_val8<SEM>(c, dst, false, tid, o);
    if constexpr (SEM == SEM_SMOOTH) {
#pragma unroll 1
      for (int j = 0; j < nsrc; ++j) {
        Vec<1> v[8];
        const bool neg = (negmask >> j) & 1;
        load_val8<SEM>(c, in.src[j], neg, tid, v);
        float gj[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float wgt = AD::pair_weight(o[i].x[0], v[i].x[0], rp);
          gj[i] = (g[i] != 0.f) ? g[i] * wgt : 0.f;
        }
        adj_accum8<SEM>(c, in.src[j], neg, (firstmask >> j) & 1, tid, gj);
      }
    } else {
      Vec<AD::KW> k[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        bool lost;
        k[i] = AD::out_coeff(g[i], o[i].x[0], sw, rp, lost);
        if (!(g[i] != 0.f)) k[i] = SumOp<AD::KW>::ident();
      }
#pragma unroll 1
      for (int j = 0; j < nsrc; ++j) {
        Vec<1> v[8];
        const bool neg = (negmask >> j) & 1;
        load_val8<SEM>(c, in.src[j], neg, tid, v);
        const float wj = in.w[j];
        float gj[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) gj[i] = (g[i] != 0.f) ? AD::in_factor(k[i], v[i].x[0], wj, rp) : 0.f;
        adj_accum8<SEM>(c, in.src[j], neg, (firstmask >> j) & 1, tid, gj);
      }
    }
  }
}

// ------------------------------------------------------------------ windows, backward
// adj(src)[q] += in_factor(x_q) * sum_{t in [q-b, q-a]} out_coeff(g_t, out_t)
template <int SEM, int MODE, class EX>
PYSTL_HD void op_win_bwd(EX& ex, const Ctx& c, const TInstr& in) {
  using AD = Adj<SEM, MODE>;
  constexpr int KW = AD::KW;
  using S = SumOp<KW>;
  RP rp = c.rp;
  Win wn = decode_win(in);
  const bool open = in.op == T_WIN_SUFFIX;
  if (open) wn.b = 0x3fffffff;
  if constexpr (SEM == SEM_SMOOTH) rp.shift = smooth_shift<MODE>(c, wn.src, wn.neg);
  // backward-looking window: t in [q - b, q - a]
  const int fa = clamp_off(wn.a, c.lext), fb = clamp_off(wn.b, c.lext);
  const int wa = -fb, wb = -fa;
  const bool vh = (wb - wa + 1) > 8;
  PYSTL_FOR_THREADS(ex, tid) {
    float g[8];
    Vec<1> o[8];
    load8(adj_plane<SEM>(c, wn.dst), tid, g);
    load_val8<SEM>(c, wn.dst, false, tid, o);
    const WinCount wk = win_count8(c, wn, 8 * tid);
    Vec<KW> k[8];
    unsigned lostmask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = wk.interior ? 1 : window_count(c, 8 * tid + i, wn.a, wn.b);
      const float sw = wk.interior ? wk.full : (float)(n > 0 ? n : 1);
      bool lost = false;
      k[i] = AD::out_coeff(g[i], o[i].x[0], sw, rp, lost);
      if (!(g[i] != 0.f) || n <= 0) {  // no cotangent / the reference raises here
        k[i] = S::ident();
        lost = false;
      }
      if (lost) lostmask |= 1u << i;
    }
    win_phase1<S, KW>(c, tid, vh, k);
    ex.put(tid, i2f((int)lostmask));
  }
  ex.sync();
  PYSTL_FOR_THREADS(ex, tid) {
    Vec<KW> w[8];
    win_phase2<S, KW>(c, tid, vh, wa, wb, w);
    Vec<1> x[8];
    load_val8<SEM>(c, wn.src, wn.neg, tid, x);
    float gx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) gx[i] = (8 * tid + i < c.nlive) ? AD::in_factor(w[i], x[i].x[0], 1.f, rp) : 0.f;
    adj_accum8<SEM>(c, wn.src, wn.neg, (in.adj_first & 1) != 0, tid, gx);
  }
  ex.sync();
  if constexpr (SEM == SEM_SMOOTH) {
    // exact scatter of the outputs whose block-shifted sum lost its precision (rare)
    PYSTL_FOR_THREADS(ex, tid) {
      const unsigned lostmask = (unsigned)f2i(ex.get(tid));
      if (lostmask) {
        float g[8];
        Vec<1> o[8];
        load8(adj_plane<SEM>(c, wn.dst), tid, g);
        load_val8<SEM>(c, wn.dst, false, tid, o);
        const float* xs = c.plane(wn.src, 0, 1);
        float* ga = adj_plane<SEM>(c, wn.src);
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          if (!((lostmask >> i) & 1)) continue;
          float gi = 0.f, oi = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j == i) {
              gi = g[j];
              oi = o[j].x[0];
            }
          const int p = 8 * tid + i;
          int lo = p + wn.a, hi = p + wn.b;
          if (hi > c.nlive - 1) hi = c.nlive - 1;
          for (int q = lo; q <= hi; ++q) {
            const float xq = wn.neg ? -xs[pidx(q)] : xs[pidx(q)];
            const float d = gi * AD::pair_weight(oi, xq, rp);
            ex.atomic_add(&ga[pidx(q)], wn.neg ? -d : d);
          }
        }
      }
    }
    ex.sync();
  }
}

// ------------------------------------------------------------------ instruction dispatch, backward
template <int SEM, class EX>
PYSTL_HD void run_instr_bwd(EX& ex, const Ctx& c, const TInstr& in) {
  if constexpr (SEM == SEM_SMOOTH || SEM == SEM_AGM || SEM == SEM_DGMSR) {
    switch (in.op) {
      case T_NARY:
        if (in.mode == MODE_AND) op_nary_bwd<SEM, MODE_AND>(ex, c, in);
        else op_nary_bwd<SEM, MODE_OR>(ex, c, in);
        break;
      case T_WIN_DIRECT:
      case T_WIN_VH:
      case T_WIN_SUFFIX:
        if (in.mode == MODE_AND) op_win_bwd<SEM, MODE_AND>(ex, c, in);
        else op_win_bwd<SEM, MODE_OR>(ex, c, in);
        break;
      default:
        break;  // T_UNTIL: rejected when the backward pass is planned
    }
  }
}

// ------------------------------------------------------------------ cotangent in, gradient out
// Root adjoint for tile positions [0, p_hi): grad_out (or the constant `scale`), zero elsewhere;
// every other adjoint plane is cleared.  go_tile = &grad_out[b, e0] or nullptr.
template <int SEM, class EX>
PYSTL_HD void init_adjoints(EX& ex, const Ctx& c, const TileProgram& P, const float* go_tile, float scale, int p_hi) {
  const int root = P.root;
  const bool neg = P.root_neg != 0;
  if (p_hi > c.last_ok + 1) p_hi = c.last_ok + 1;
  PYSTL_FOR_THREADS(ex, tid) {
    float g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = 8 * tid + i;
      float v = 0.f;
      if (p < p_hi) v = go_tile ? go_tile[p] : scale;
      g[i] = neg ? -v : v;
    }
    store8(adj_plane<SEM>(c, root), tid, g);
  }
}

// copy the adjoint-plane map of the program next to the planes (uniform reads afterwards)
template <class EX>
PYSTL_HD void load_adj_table(EX& ex, const Ctx& c, const TileProgram& P) {
  PYSTL_FOR_THREADS(ex, tid) {
    for (int s = tid; s < kMaxSlots; s += c.nt) c.adj_tab()[s] = P.adj_of[s];
  }
  ex.sync();
}

// grad_signal[p, d] = sum over predicates of coeff[d] * adjoint(predicate)[p] for tile positions
// [p_lo, p_hi); every channel is written (zeros for unreferenced ones).  gs_tile = &grad_signal[b, e0, 0].
template <int SEM, class EX>
PYSTL_HD void store_signal_grad(EX& ex, const Ctx& c, const TileProgram& P, float* gs_tile, int p_lo, int p_hi) {
  const int D = c.D, nt = c.nt, np = P.n_preds;
  ex.sync();
  PYSTL_FOR_THREADS(ex, tid) {
#pragma unroll 1
    for (int i = 0; i < kChunk; ++i) {
      const int p = tid + nt * i;  // strided ownership: coalesced global stores
      if (p < p_lo || p >= p_hi) continue;
      const int pp = pidx(p);
      if (D <= 8) {
        float row[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) row[d] = 0.f;
#pragma unroll 1
        for (int q = 0; q < np; ++q) {
          const float g = c.adj_tab()[P.preds[q].dst] >= 0 ? adj_plane<SEM>(c, P.preds[q].dst)[pp] : 0.f;
          const int nterms = P.preds[q].n_terms;
#pragma unroll 1
          for (int k = 0; k < nterms; ++k) {
            const int ch = P.preds[q].ch[k];
            const float cg = P.preds[q].coeff[k] * g;
#pragma unroll
            for (int d = 0; d < 8; ++d) row[d] += (ch == d) ? cg : 0.f;
          }
        }
#pragma unroll
        for (int d = 0; d < 8; ++d)
          if (d < D) gs_tile[(int64_t)p * D + d] = row[d];
      } else {
        for (int d = 0; d < D; ++d) {
          float acc = 0.f;
          for (int q = 0; q < np; ++q) {
            const float g = c.adj_tab()[P.preds[q].dst] >= 0 ? adj_plane<SEM>(c, P.preds[q].dst)[pp] : 0.f;
            for (int k = 0; k < P.preds[q].n_terms; ++k)
              if (P.preds[q].ch[k] == d) acc += P.preds[q].coeff[k] * g;
          }
          gs_tile[(int64_t)p * D + d] = acc;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ whole backward pass of one tile
// p_lo / p_hi: tile positions of the samples this CTA owns the gradient of.
template <int SEM, class EX>
PYSTL_HD void run_backward(EX& ex, const Ctx& c, const TileProgram& P, const float* raw, const float* go_tile,
                           float scale, float* gs_tile, int p_lo, int p_hi) {
  load_adj_table(ex, c, P);
  if constexpr (SEM == SEM_CLASSIC) {
    constexpr int S = SEM_CLASSIC_SRC;
    run_program<S>(ex, c, P, raw);
    ex.sync();
    // adjoint planes of the predicates start at zero
    PYSTL_FOR_THREADS(ex, tid) {
      float z[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) z[i] = 0.f;
      for (int q = 0; q < P.n_preds; ++q) store8(adj_plane<S>(c, P.preds[q].dst), tid, z);
    }
    ex.sync();
    // one scatter of the root cotangent through the recorded leaf
    int hi = p_hi;
    if (hi > c.last_ok + 1) hi = c.last_ok + 1;
    const int root = P.root;
    const bool rneg = P.root_neg != 0;
    PYSTL_FOR_THREADS(ex, tid) {
      float src[8];
      load8(c.plane(root, 1, 2), tid, src);
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        const int p = 8 * tid + i;
        if (p >= hi) continue;
        float sv = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j == i) sv = src[j];
        const float g = go_tile ? go_tile[p] : scale;
        if (!(g != 0.f)) continue;
        const int s = f2i(sv);
        const int q = (s >> 24) & 0x7f, pos = s & 0xffffff;
        const bool neg = (s < 0) != rneg;
        ex.atomic_add(&adj_plane<S>(c, P.preds[q].dst)[pidx(pos)], neg ? -g : g);
      }
    }
    store_signal_grad<S>(ex, c, P, gs_tile, p_lo, p_hi);
  } else {
    run_program<SEM>(ex, c, P, raw);
    ex.sync();
    init_adjoints<SEM>(ex, c, P, go_tile, scale, p_hi);
    // own-chunk accumulation everywhere: no barrier needed between element-wise instructions
#pragma unroll 1
    for (int i = P.n_instr - 1; i >= 0; --i) run_instr_bwd<SEM>(ex, c, P.ins[i]);
    store_signal_grad<SEM>(ex, c, P, gs_tile, p_lo, p_hi);
  }
}

}  // namespace pystl
