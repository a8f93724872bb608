// Lowering: pystl_formula_t (post-order node list) -> TileProgram + launch geometry.
// Host-only C++, shared by the CUDA launcher and the test-only host emulator.
//
// What is decided here
//   * validity: interval / weight checks with the reference's error classes
//     (api.py:36-40, agm.py:29-34), start_slack (empty-window ValueError region,
//     api.py:218-221) and the "until child with a start raises everywhere" quirk
//     (api.py:288-293);
//   * Not is never an instruction: Not(predicate) becomes a predicate with negated
//     coefficients (exact in IEEE arithmetic), any other Not becomes a negate bit on
//     the consumer's operand;
//   * identical predicates share a plane;
//   * bounded windows pick the direct (W <= 8) or van Herk (W >= 9) method;
//   * classic until with a long window is decomposed into two windows and a clamp
//     scan; everything else uses the O(b) until;
//   * plane slots are reference-counted and reused.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/pystl_b200.h"
#include "program.h"

namespace pystl {
namespace {

constexpr int kNever = 1 << 29;  // start slack of a node that raises for every t

struct Fail {
  int code;
};

struct Lowerer {
  const pystl_formula_t* f = nullptr;
  bool backward = false;
  bool ssa = false;  // fused tier: every value keeps its own slot
  Lowered* out = nullptr;
  char* msg = nullptr;
  int msg_len = 0;

  struct Ref {
    int slot = -1;
    bool neg = false;
  };
  std::vector<int> look;       // lookahead per node, -1 unbounded
  std::vector<int> slack;      // start slack per node
  std::vector<int> slot_refs;  // outstanding readers per slot
  std::vector<int> free_slots;
  int n_slots = 0;

  [[noreturn]] void fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    if (msg && msg_len > 0) vsnprintf(msg, msg_len, fmt, ap);
    va_end(ap);
    throw Fail{code};
  }

  int alloc_slot() {
    int s;
    if (!free_slots.empty()) {
      s = free_slots.back();
      free_slots.pop_back();
    } else {
      s = n_slots++;
      slot_refs.push_back(0);
    }
    slot_refs[s] = 0;
    return s;
  }
  void release(int slot) {
    if (--slot_refs[slot] == 0 && !backward && !ssa) free_slots.push_back(slot);  // backward / ssa: no reuse
  }

  TileProgram& P() { return out->prog; }

  int find_or_add_pred(const pystl_pred_t& src, bool negate) {
    TPred p;
    std::memset(&p, 0, sizeof(p));
    p.n_terms = (int16_t)src.n_terms;
    for (int k = 0; k < src.n_terms; ++k) {
      p.ch[k] = (int16_t)src.channel[k];
      p.coeff[k] = negate ? -src.coeff[k] : src.coeff[k];
    }
    p.bias = negate ? -src.bias : src.bias;
    p.pair = src.pair ? 1 : 0;  // (never negated: resolve() keeps the NOT above a pair)
    // +0.0 / -0.0 biases are the same predicate
    if (p.bias == 0.f) p.bias = 0.f;
    for (int q = 0; q < P().n_preds; ++q) {
      const TPred& o = P().preds[q];
      if (o.n_terms != p.n_terms || o.bias != p.bias || o.pair != p.pair) continue;  // (f64 is a function of these)
      bool same = true;
      for (int k = 0; k < p.n_terms && same; ++k) same = o.ch[k] == p.ch[k] && o.coeff[k] == p.coeff[k];
      if (same) return q;
    }
    if (P().n_preds >= kMaxPreds) fail(PYSTL_E_UNSUPPORTED, "more than %d distinct predicates", kMaxPreds);
    // fp32 is already bit-identical to float32(float64 evaluation) when a single rounding
    // happens: +-x + bias, or +-x +- y.  Anything else is evaluated in float64 for classic.
    bool unit = true;
    for (int k = 0; k < p.n_terms; ++k) unit = unit && (p.coeff[k] == 1.f || p.coeff[k] == -1.f);
    const bool single_rounding = unit && (p.n_terms <= 1 || (p.n_terms == 2 && p.bias == 0.f));
    p.f64 = (f->semantics == SEM_CLASSIC && !single_rounding) ? 1 : 0;
    p.dst = (int16_t)alloc_slot();
    P().preds[P().n_preds] = p;
    return P().n_preds++;
  }

  TInstr& emit(int op, int mode) {
    if (P().n_instr >= kMaxInstr) fail(PYSTL_E_UNSUPPORTED, "formula lowers to more than %d instructions", kMaxInstr);
    TInstr& in = P().ins[P().n_instr++];
    std::memset(&in, 0, sizeof(in));
    in.op = (int16_t)op;
    in.mode = (int16_t)mode;
    in.w_off = in.w2_off = -1;
    return in;
  }
  void add_src(TInstr& in, Ref r) {
    in.src[in.nsrc] = (int16_t)r.slot;
    if (r.neg) in.negmask |= (uint16_t)(1u << in.nsrc);
    in.w[in.nsrc] = 1.f;
    ++in.nsrc;
  }

  // window [a, b] (b = -1: open) over `src`
  Ref window(Ref src, int mode, int a, int b) {
    const int dst = alloc_slot();
    int op;
    if (b < 0) op = T_WIN_SUFFIX;
    else op = (b - a + 1 <= 8) ? T_WIN_DIRECT : T_WIN_VH;
    TInstr& in = emit(op, mode);
    add_src(in, src);
    in.dst = (int16_t)dst;
    in.a = a;
    in.b = b < 0 ? 0 : b;
    in.unbounded = b < 0;
    return Ref{dst, false};
  }

  const float* weights_of(int off, int len, int need, const char* what, bool agm_rules) {
    if (off < 0) return nullptr;
    if (off + len > f->n_weights || len < 0) fail(PYSTL_E_INVALID, "%s: weight range out of bounds", what);
    if (len < need) fail(PYSTL_E_INVALID, "%s requires at least %d entries.", what, need);  // agm.py:29-30
    const float* w = f->weights + off;
    if (agm_rules) {  // agm.py:31-34
      double s = 0;
      for (int i = 0; i < need; ++i) {
        if (w[i] < 0.f) fail(PYSTL_E_INVALID, "%s entries must be >= 0.", what);
        s += w[i];
      }
      if (s <= 0) fail(PYSTL_E_INVALID, "%s entries must contain at least one positive weight.", what);
    }
    return w;
  }

  // copy `n` window-position weights into the program (fused tier); returns their offset
  int stash_weights(const float* w, int n) {
    if (P().n_wbuf + n > kMaxWeights) fail(PYSTL_E_UNSUPPORTED, "more than %d window weights in one formula", kMaxWeights);
    const int off = P().n_wbuf;
    for (int i = 0; i < n; ++i) P().wbuf[off + i] = w[i];
    P().n_wbuf += n;
    return off;
  }

  // ---- structural resolution: which plane (and sign) does an operand read? ----
  struct Val {
    int owner = -1;  // predicate index or node index
    bool is_pred = false;
    bool neg = false;
  };
  std::vector<Val> val;
  std::vector<char> val_done;
  int child(int i, int k) const { return f->children[f->nodes[i].child_off + k]; }

  Val resolve(int i) {
    if (val_done[i]) return val[i];
    const pystl_node_t& nd = f->nodes[i];
    Val v;
    if (nd.kind == PYSTL_NODE_PRED) {
      v.owner = find_or_add_pred(f->preds[nd.pred], false);
      v.is_pred = true;
    } else if (nd.kind == PYSTL_NODE_NOT) {
      const int ch = child(i, 0);
      if (f->nodes[ch].kind == PYSTL_NODE_PRED && !f->preds[f->nodes[ch].pred].pair) {  // fold the negation into the coefficients
        v.owner = find_or_add_pred(f->preds[f->nodes[ch].pred], true);
        v.is_pred = true;
      } else {
        v = resolve(ch);
        v.neg = !v.neg;
      }
    } else {
      v.owner = i;
    }
    val[i] = v;
    val_done[i] = 1;
    return v;
  }

  void run() {
    const int sem = f->semantics;
    if (sem < SEM_CLASSIC || sem > SEM_DGMSR) fail(PYSTL_E_INVALID, "unknown semantics %d", sem);
    if (f->n_nodes <= 0 || !f->nodes) fail(PYSTL_E_INVALID, "formula has no nodes");
    std::memset(&P(), 0, sizeof(TileProgram));
    P().sem = sem;
    P().tau = f->temperature;
    P().eps = f->eps;
    P().p = f->p;
    {
      RP& r = P().rp;
      const double tau = f->temperature > 0.f ? f->temperature : 1.0;
      const int pp = f->p < 1 ? 1 : f->p;
      const double eps = f->eps > 0.f ? f->eps : 1e-8;
      r.k = (float)(1.4426950408889634 / tau);
      r.invk = (float)(tau * 0.6931471805599453);
      r.shift = 0.f;
      r.eps = (float)eps;
      r.sqrt_eps = (float)std::sqrt(eps);
      r.ln_eps = (float)std::log(eps);
      r.eps_p = (float)std::pow(eps, (double)pp);
      r.inv2p = (float)(0.5 / pp);
      r.p = pp;
    }
    if (backward && sem == SEM_CUMULATIVE && !ssa)
      fail(PYSTL_E_UNSUPPORTED, "the gradient of the cumulative semantics needs the fused tier (NVRTC)");
    if (sem == SEM_SMOOTH && !(f->temperature > 0.f)) fail(PYSTL_E_INVALID, "temperature must be > 0.");  // smooth.py:12-15
    if (sem == SEM_DGMSR && (!(f->eps > 0.f) || f->p < 1)) fail(PYSTL_E_INVALID, "dgmsr needs eps > 0 and p >= 1");
    const bool uses_weights = (sem == SEM_AGM || sem == SEM_DGMSR);
    const int n = f->n_nodes;
    look.assign(n, 0);
    slack.assign(n, 0);
    val.assign(n, Val{});
    val_done.assign(n, 0);

    // ---- validation of the node list; look-ahead and start slack per node ----
    for (int i = 0; i < n; ++i) {
      const pystl_node_t& nd = f->nodes[i];
      if (nd.n_children < 0 || nd.child_off < 0 || nd.child_off + nd.n_children > f->n_children)
        fail(PYSTL_E_INVALID, "node %d: children out of range", i);
      for (int k = 0; k < nd.n_children; ++k) {
        const int ch = child(i, k);
        if (ch < 0 || ch >= i) fail(PYSTL_E_INVALID, "node %d: nodes must be in post-order", i);
      }
      auto max_look = [&]() {
        int lk = 0;
        for (int k = 0; k < nd.n_children; ++k) lk = (lk < 0 || look[child(i, k)] < 0) ? -1 : std::max(lk, look[child(i, k)]);
        return lk;
      };
      auto max_slack = [&]() {
        int sl = 0;
        for (int k = 0; k < nd.n_children; ++k) sl = std::max(sl, slack[child(i, k)]);
        return sl;
      };
      switch (nd.kind) {
        case PYSTL_NODE_PRED: {
          if (nd.pred < 0 || nd.pred >= f->n_preds) fail(PYSTL_E_INVALID, "node %d: predicate index out of range", i);
          const pystl_pred_t& pr = f->preds[nd.pred];
          if (pr.n_terms < 0 || pr.n_terms > kMaxTerms) fail(PYSTL_E_UNSUPPORTED, "predicate with more than %d terms", kMaxTerms);
          if (pr.pair) {
            if (sem != SEM_CUMULATIVE) fail(PYSTL_E_INVALID, "node %d: a pair predicate is a cumulative value", i);
            if (pr.n_terms != 2) fail(PYSTL_E_INVALID, "node %d: a pair predicate names two channels (pos, neg)", i);
            if (!ssa) fail(PYSTL_E_UNSUPPORTED, "pair predicates (host-computed cumulative values) need the fused tier (NVRTC)");
          }
          // a channel of values the host computed (open-ended window beyond one tile) that end before the horizon
          slack[i] = pr.slack < 0 ? kNever : std::min(pr.slack, kNever);
          break;
        }
        case PYSTL_NODE_NOT:
          if (nd.n_children != 1) fail(PYSTL_E_INVALID, "node %d: NOT takes one child", i);
          look[i] = max_look();
          slack[i] = max_slack();
          break;
        case PYSTL_NODE_AND:
        case PYSTL_NODE_OR:
          if (nd.n_children < 1) fail(PYSTL_E_INVALID, "node %d: And/Or requires at least one child.", i);  // api.py:166-167
          if (nd.n_children > kMaxSrc) fail(PYSTL_E_UNSUPPORTED, "And/Or with more than %d children", kMaxSrc);
          look[i] = max_look();
          slack[i] = max_slack();
          break;
        case PYSTL_NODE_ALWAYS:
        case PYSTL_NODE_EVENTUALLY:
        case PYSTL_NODE_UNTIL: {
          const int want = nd.kind == PYSTL_NODE_UNTIL ? 2 : 1;
          if (nd.n_children != want) fail(PYSTL_E_INVALID, "node %d: temporal operator takes %d operand(s)", i, want);
          if (nd.a < 0) fail(PYSTL_E_INVALID, "Interval start must be >= 0.");                                 // api.py:37-38
          if (nd.b < -1 || (nd.b >= 0 && nd.b < nd.a)) fail(PYSTL_E_INVALID, "Interval end must be >= start.");  // api.py:39-40
          const int lk = max_look();
          look[i] = (nd.b < 0 || lk < 0) ? -1 : nd.b + lk;
          // Last t that does not raise is T-1-slack.  A child that is valid everywhere only
          // needs a non-empty window (t+a <= T-1, api.py:218-221).  A child with its own
          // slack raises inside any window that is clipped at the horizon, so the whole
          // window must fit: t+b <= T-1-slack_child.  UNTIL evaluates its children up to the
          // horizon (api.py:288-293), an open window reaches it: both then raise for every t.
          const int cs = max_slack();
          if (cs == 0) slack[i] = nd.a;
          else if (nd.kind == PYSTL_NODE_UNTIL || nd.b < 0 || cs >= kNever) slack[i] = kNever;
          else slack[i] = std::min(kNever, cs + nd.b);
          break;
        }
        default:
          fail(PYSTL_E_INVALID, "node %d: unknown kind %d", i, nd.kind);
      }
    }

    // ---- reachability and read counts (Not is transparent) ----
    std::vector<char> reach(n, 0);
    reach[n - 1] = 1;
    for (int i = n - 1; i >= 0; --i)
      if (reach[i])
        for (int k = 0; k < f->nodes[i].n_children; ++k) reach[child(i, k)] = 1;
    auto emits = [&](int i) {
      const int k = f->nodes[i].kind;
      return k != PYSTL_NODE_PRED && k != PYSTL_NODE_NOT;
    };
    std::vector<int> node_uses(n, 0), pred_uses(kMaxPreds, 0);
    auto count_use = [&](Val v) { (v.is_pred ? pred_uses[v.owner] : node_uses[v.owner]) += 1; };
    for (int i = 0; i < n; ++i)
      if (reach[i] && emits(i))
        for (int k = 0; k < f->nodes[i].n_children; ++k) count_use(resolve(child(i, k)));
    const Val root = resolve(n - 1);
    count_use(root);
    // predicate planes are written before the first instruction: they own the first slots
    for (int q = 0; q < P().n_preds; ++q) slot_refs[P().preds[q].dst] = pred_uses[q];

    std::vector<int> node_slot(n, -1);
    auto operand = [&](int node) -> Ref {
      const Val v = resolve(node);
      return Ref{v.is_pred ? P().preds[v.owner].dst : node_slot[v.owner], v.neg};
    };

    // ---- instructions ----
    for (int i = 0; i < n; ++i) {
      if (!reach[i] || !emits(i)) continue;
      const pystl_node_t& nd = f->nodes[i];
      Ref res;
      switch (nd.kind) {
        case PYSTL_NODE_AND:
        case PYSTL_NODE_OR: {
          const float* w = uses_weights ? weights_of(nd.w_off, nd.w_len, nd.n_children, "weights", sem == SEM_AGM) : nullptr;
          const int dst = alloc_slot();
          TInstr& in = emit(T_NARY, nd.kind == PYSTL_NODE_AND ? MODE_AND : MODE_OR);
          double sw = 0;
          for (int k = 0; k < nd.n_children; ++k) {
            add_src(in, operand(child(i, k)));
            if (w) in.w[k] = w[k];
            sw += in.w[k];
          }
          in.sw = (float)sw;
          in.dst = (int16_t)dst;
          res = Ref{dst, false};
          break;
        }
        case PYSTL_NODE_ALWAYS:
        case PYSTL_NODE_EVENTUALLY: {
          int mode = nd.kind == PYSTL_NODE_ALWAYS ? MODE_AND : MODE_OR;
          if (sem == SEM_CUMULATIVE && nd.kind == PYSTL_NODE_EVENTUALLY) mode = MODE_SUM;  // cumulative.py:61-72
          if (uses_weights && nd.w_off >= 0) {
            // per-position weights, truncated to the clipped window (agm.py:28-35, dgmsr.py:20-25)
            if (!ssa) fail(PYSTL_E_UNSUPPORTED, "window weights on Always / Eventually need the fused tier (NVRTC)");
            if (nd.b < 0) fail(PYSTL_E_UNSUPPORTED, "window weights on an open-ended interval");
            const int W = nd.b - nd.a + 1;
            const float* w = weights_of(nd.w_off, nd.w_len, W, "weights", sem == SEM_AGM);
            const Ref src = operand(child(i, 0));
            const int dst = alloc_slot();
            TInstr& in = emit(T_WIN_WEIGHTED, mode);
            add_src(in, src);
            in.dst = (int16_t)dst;
            in.a = nd.a;
            in.b = nd.b;
            in.w_off = (int16_t)stash_weights(w, W);
            res = Ref{dst, false};
            break;
          }
          res = window(operand(child(i, 0)), mode, nd.a, nd.b);
          break;
        }
        case PYSTL_NODE_UNTIL: {
          if (backward && sem != SEM_CLASSIC && !ssa)
            fail(PYSTL_E_UNSUPPORTED, "the gradient of a non-classical Until needs the fused tier (NVRTC)");
          const bool until_weights = uses_weights && (nd.w_off >= 0 || nd.w2_off >= 0);
          if (until_weights && (!ssa || nd.b < 0))
            fail(PYSTL_E_UNSUPPORTED, "weights_left / weights_right on Until need the fused tier (bounded interval)");
          float wp0 = 1.f, wp1 = 1.f;
          if (uses_weights) {
            wp0 = nd.wp0;
            wp1 = nd.wp1;
            if (sem == SEM_AGM) {  // agm.py:31-34 on weights_pair
              if (wp0 < 0.f || wp1 < 0.f) fail(PYSTL_E_INVALID, "weights_pair entries must be >= 0.");
              if (wp0 + wp1 <= 0.f) fail(PYSTL_E_INVALID, "weights_pair entries must contain at least one positive weight.");
            }
          }
          const Ref l = operand(child(i, 0)), r = operand(child(i, 1));
          // (the fused tier always decomposes: every piece is O(1) per sample and fusable)
          const bool decompose = sem == SEM_CLASSIC && (nd.b < 0 || nd.b > 16 || ssa);
          if (decompose) {
            // U[a,b] = min( G[0,a] l , W(t+a) , F[a,b] r ),  W = unbounded until (clamp scan)
            // (every slot below is held by one pending read until the final min consumes it)
            const int s_scan = alloc_slot();
            {
              TInstr& in = emit(T_CLAMP_SCAN, MODE_AND);
              add_src(in, l);
              add_src(in, r);
              in.dst = (int16_t)s_scan;
            }
            slot_refs[s_scan] = 1;
            Ref w_shift{s_scan, false}, g_l{-1, false}, f_r{-1, false};
            if (nd.a > 0) {
              w_shift = window(Ref{s_scan, false}, MODE_AND, nd.a, nd.a);
              slot_refs[w_shift.slot] = 1;
              release(s_scan);
              g_l = window(l, MODE_AND, 0, nd.a);
              slot_refs[g_l.slot] = 1;
            }
            if (nd.b >= 0) {
              f_r = window(r, MODE_OR, nd.a, nd.b);
              slot_refs[f_r.slot] = 1;
            }
            const int dst = alloc_slot();
            TInstr& in = emit(T_NARY, MODE_AND);
            add_src(in, w_shift);
            if (nd.a > 0) add_src(in, g_l);
            if (nd.b >= 0) add_src(in, f_r);
            in.sw = (float)in.nsrc;
            in.dst = (int16_t)dst;
            release(w_shift.slot);
            if (nd.a > 0) release(g_l.slot);
            if (nd.b >= 0) release(f_r.slot);
            res = Ref{dst, false};
          } else {
            const int dst = alloc_slot();
            TInstr& in = emit(T_UNTIL, MODE_AND);
            add_src(in, l);
            add_src(in, r);
            in.w[0] = wp0;
            in.w[1] = wp1;
            in.a = nd.a;
            in.b = nd.b < 0 ? 0 : nd.b;
            in.unbounded = nd.b < 0;
            in.dst = (int16_t)dst;
            if (until_weights) {
              // weights_left[:i+1] over the prefix, weights_right[:n] over the candidates (agm.py:112-130)
              if (nd.w_off >= 0) {
                const float* wl = weights_of(nd.w_off, nd.w_len, nd.b + 1, "weights_left", sem == SEM_AGM);
                P().ins[P().n_instr - 1].w_off = (int16_t)stash_weights(wl, nd.b + 1);
              }
              if (nd.w2_off >= 0) {
                const float* wr = weights_of(nd.w2_off, nd.w2_len, nd.b - nd.a + 1, "weights_right", sem == SEM_AGM);
                P().ins[P().n_instr - 1].w2_off = (int16_t)stash_weights(wr, nd.b - nd.a + 1);
              }
            }
            res = Ref{dst, false};
          }
          break;
        }
        default:
          break;
      }
      node_slot[i] = res.slot;
      slot_refs[res.slot] = node_uses[i];
      // the operands have been read once each
      for (int k = 0; k < nd.n_children; ++k) {
        const Ref o = operand(child(i, k));
        release(o.slot);
      }
    }
    const Ref rr = operand(n - 1);
    if (n_slots > kMaxSlots) fail(PYSTL_E_UNSUPPORTED, "formula needs more than %d value planes", kMaxSlots);
    P().root = (int16_t)rr.slot;
    P().root_neg = (int16_t)(rr.neg ? 1 : 0);
    P().n_slots = n_slots;
    P().start_slack = slack[n - 1] >= kNever ? 0 : slack[n - 1];
    P().raises_always = slack[n - 1] >= kNever ? 1 : 0;
    out->lookahead = look[n - 1];
    out->out_channels = sem == SEM_CUMULATIVE ? 2 : 1;
    if (backward) plan_adjoint_planes();
  }

  // Backward pass: assign a plane to every adjoint, reusing planes whose contents are dead.
  // Instructions run in reverse; after instruction i its result's value and adjoint are dead
  // (every consumer has a higher index and has already run), and an operand's value dies
  // with its lowest-index consumer.  Classic only needs the predicates' adjoints (the
  // gradient is one scatter through the recorded provenance).
  void plan_adjoint_planes() {
    TileProgram& p = P();
    const int vw = p.sem == SEM_CLASSIC ? 2 : 1;  // SEM_CLASSIC_SRC values are two words
    for (int s = 0; s < kMaxSlots; ++s) p.adj_of[s] = -1;
    int n_planes = p.n_slots * vw;
    if (p.sem == SEM_CLASSIC) {
      for (int q = 0; q < p.n_preds; ++q) p.adj_of[p.preds[q].dst] = (int16_t)n_planes++;
      p.n_planes_bwd = n_planes;
      return;
    }
    std::vector<int> free_planes;
    auto take = [&]() {
      if (!free_planes.empty()) {
        const int x = free_planes.back();
        free_planes.pop_back();
        return x;
      }
      return n_planes++;
    };
    std::vector<int> first_use(p.n_slots, p.n_instr);  // lowest-index instruction reading the value
    std::vector<char> is_pred(p.n_slots, 0);
    for (int q = 0; q < p.n_preds; ++q) is_pred[p.preds[q].dst] = 1;
    for (int i = 0; i < p.n_instr; ++i)
      for (int j = 0; j < p.ins[i].nsrc; ++j) first_use[p.ins[i].src[j]] = std::min(first_use[p.ins[i].src[j]], i);
    p.adj_of[p.root] = (int16_t)take();  // the cotangent lands here
    for (int i = p.n_instr - 1; i >= 0; --i) {
      TInstr& in = p.ins[i];
      in.adj_first = 0;
      if (p.adj_of[in.dst] < 0) continue;  // result never reaches the root (cannot happen: dead code is not lowered)
      for (int j = 0; j < in.nsrc; ++j) {
        if (p.adj_of[in.src[j]] < 0) {
          p.adj_of[in.src[j]] = (int16_t)take();
          in.adj_first |= (uint16_t)(1u << j);
        }
      }
      // the result's adjoint and value are dead now; so is any operand value this was the last reader of
      free_planes.push_back(p.adj_of[in.dst]);
      free_planes.push_back(in.dst);
      for (int j = 0; j < in.nsrc; ++j) {
        const int sj = in.src[j];
        bool again = false;
        for (int k = 0; k < j; ++k) again = again || in.src[k] == sj;
        if (!again && first_use[sj] == i && is_pred[sj]) free_planes.push_back(sj);
      }
    }
    p.n_planes_bwd = n_planes;
  }
};

}  // namespace

int lower_formula(const ::pystl_formula* f, bool backward, Lowered* out, char* msg, int msg_len) {
  return lower_formula_ex(f, backward, false, out, msg, msg_len);
}

int lower_formula_ex(const ::pystl_formula* f, bool backward, bool ssa, Lowered* out, char* msg, int msg_len) {
  if (msg && msg_len > 0) msg[0] = 0;
  if (!f || !out) {
    if (msg) snprintf(msg, msg_len, "null formula");
    return PYSTL_E_INVALID;
  }
  Lowerer L;
  L.f = reinterpret_cast<const pystl_formula_t*>(f);
  L.backward = backward;
  L.ssa = ssa;
  L.out = out;
  L.msg = msg;
  L.msg_len = msg_len;
  try {
    L.run();
  } catch (const Fail& e) {
    return e.code;
  }
  return PYSTL_OK;
}

static bool needs_full_trace(const TileProgram& P) {
  for (int i = 0; i < P.n_instr; ++i)
    if (P.ins[i].unbounded) return true;
  return false;
}

int plan_geometry(const Lowered& lw, int64_t B, int64_t T, int32_t D, int force_nt, bool backward, TileGeom* g, char* msg,
                  int msg_len) {
  const TileProgram& P = lw.prog;
  auto fail = [&](int code, const char* fmt, auto... a) {
    if (msg && msg_len > 0) snprintf(msg, msg_len, fmt, a...);
    return code;
  };
  if (B < 0 || T < 0 || D < 1) return fail(PYSTL_E_INVALID, "bad shape B=%lld T=%lld D=%d", (long long)B, (long long)T, D);
  for (int q = 0; q < P.n_preds; ++q)
    for (int k = 0; k < P.preds[q].n_terms; ++k)
      if (P.preds[q].ch[k] < 0 || P.preds[q].ch[k] >= D)
        return fail(PYSTL_E_INVALID, "predicate reads channel %d of a %d-channel signal", (int)P.preds[q].ch[k], D);
  // the classic backward kernel evaluates SEM_CLASSIC_SRC: two words per value
  const int vw = (P.sem == SEM_CUMULATIVE || (backward && P.sem == SEM_CLASSIC)) ? 2 : 1;
  const int aw = (P.sem == SEM_CLASSIC || P.sem == SEM_SMOOTH) ? 1 : 2;
  // scratch planes: SUF + PRE of the widest accumulator (forward: aw words; the adjoint window
  // sums of the backward pass: 1 word for smooth, 2 for agm / dgmsr), or the staged until operands
  const int kw = P.sem == SEM_SMOOTH ? 1 : 2;
  const int scr_planes = backward ? 2 * std::max(aw, P.sem == SEM_CLASSIC ? 2 : kw) : 2 * std::max(aw, vw);
  const int hh = backward && lw.lookahead > 0 ? ((lw.lookahead + 7) / 8) * 8 : 0;
  const bool full = needs_full_trace(P) || lw.lookahead < 0;
  // furthest offsets any window instruction reads at (guards of the scratch planes)
  int max_b = 0, min_a = 0;
  for (int i = 0; i < P.n_instr; ++i) {
    const TInstr& in = P.ins[i];
    if (in.op == T_WIN_DIRECT || in.op == T_WIN_VH || in.op == T_WIN_SUFFIX) {
      max_b = std::max(max_b, std::max(in.a, in.b));
      min_a = std::min(min_a, std::min(in.a, in.b));
    }
  }
  const int64_t Tn = std::max<int64_t>(T, 1);
  const int smem_cap = 227 * 1024;
  // geometry for one CTA size; false if it does not fit
  auto build = [&](int nt, TileGeom* out) -> bool {
    TileGeom t{};
    t.B = B;
    t.T = T;
    t.D = D;
    t.nt = nt;
    t.lext = kChunk * nt;
    t.vw = vw;
    t.scr_planes = scr_planes;
    t.bwd = backward ? 1 : 0;
    t.hh = hh;
    t.n_slots = P.n_slots;
    if (Tn <= t.lext) {
      t.lout = t.lext;
      t.n_tiles = 1;
    } else {
      if (full) return false;
      t.lout = backward ? t.lext - 2 * hh : ((t.lext - lw.lookahead) / 8) * 8;
      if (t.lout < 8) return false;
      t.n_tiles = (int)((Tn + t.lout - 1) / t.lout);
    }
    t.gr = ((std::min(max_b, t.lext) + 4 + 7) / 8) * 8;
    t.gl = ((std::min(-min_a, t.lext) + 7) / 8) * 8;
    if (backward) {  // the adjoint windows look back as far as the forward ones look ahead
      const int reach = needs_full_trace(P) ? t.lext : std::min(std::max(max_b, -min_a), t.lext);
      t.gl = t.gr = ((reach + 8 + 7) / 8) * 8;
    }
    t.off_planes = 0;
    t.off_scratch = (backward ? P.n_planes_bwd : P.n_slots * vw) * t.lext;
    // the raw [sample][channel] rows share their storage with the scratch planes
    const int raw_floats = ((t.lext * D + 3) / 4) * 4;
    t.off_chunk = t.off_scratch + std::max(raw_floats, scr_planes * (t.gl + t.lext + t.gr));
    t.off_red = t.off_chunk + 4 * (t.gl / 8 + nt + t.gr / 8);
    t.smem_floats = t.off_red + 128 + 2 * (P.n_slots + 1) + kMaxSlots;  // reduction scratch, value ranges, adjoint map
    if (t.smem_floats * 4 > smem_cap) return false;
    *out = t;
    return true;
  };
  if (force_nt > 0) {  // tests: exercise the multi-tile path on short traces
    if (force_nt % 8 != 0 || force_nt > kMaxThreads) return fail(PYSTL_E_INVALID, "forced CTA size %d", force_nt);
    if (!build(force_nt, g)) return fail(PYSTL_E_UNSUPPORTED, "forced CTA size %d does not fit this formula", force_nt);
    return PYSTL_OK;
  }
  // CTA sizes: every multiple of 32 up to 512 threads (tiles of 256 .. 4096 samples), then 768 / 1024
  int cand[32], n_cand = 0;
  for (int nt = 512; nt >= 64; nt -= 32) cand[n_cand++] = nt;  // ties go to the larger tile
  cand[n_cand++] = 768;
  cand[n_cand++] = 1024;
  double best_cost = 0;
  int best = -1;
  TileGeom bg{};
  for (int ci = 0; ci < n_cand; ++ci) {
    TileGeom t{};
    if (!build(cand[ci], &t)) continue;
    if (t.n_tiles > 1 && t.lout < t.lext / 4) continue;  // more than 3/4 of every tile would be halo
    // Cost = samples staged per trace (halo + tail waste), scaled by how well the CTA shape
    // fills an SM.  All kernels are built for 64 registers per thread (2048 threads / SM), so
    // residency is limited by shared memory; measured on B200, throughput keeps improving up
    // to ~32 resident warps and larger tiles amortise the per-tile fixed costs.
    double cost = (double)t.n_tiles * t.lext;
    const int smem_bytes = t.smem_floats * 4 + 1024;
    int ctas = std::min((int)(227 * 1024 / smem_bytes), 2048 / t.nt);
    if (ctas < 1) ctas = 1;
    const int warps = ctas * t.nt / 32;
    if (warps < 32) cost *= 1.0 + 0.6 * (32 - warps) / 32.0;
    if (t.nt < 256) cost *= 1.0 + 0.15 * (256 - t.nt) / 256.0;  // per-tile fixed costs
    if (best < 0 || cost < best_cost * 0.98) {
      best = ci;
      best_cost = cost;
      bg = t;
    }
  }
  if (best < 0) {
    if (full)
      return fail(PYSTL_E_UNSUPPORTED,
                  "open-ended intervals need the whole trace in one tile: T=%lld exceeds what fits in shared memory",
                  (long long)T);
    return fail(PYSTL_E_UNSUPPORTED, "formula look-ahead %d / plane count %d does not fit a shared-memory tile",
                lw.lookahead, P.n_slots);
  }
  *g = bg;
  return PYSTL_OK;
}

}  // namespace pystl
