// Code generator of the FUSED tier: lowered formula (SSA TileProgram) -> CUDA kernel source.
//
// The generated kernel is glue: a straight-line sequence of calls into the hand-written
// templates of fused_rt.cuh with every structural constant (window offsets, operand signs,
// channel indices, shared-memory offsets, CTA size) baked in as a template argument, and every
// numeric parameter (predicate coefficients, weights, temperature, eps) read from the
// TileProgram the kernel receives by value -- retuning a threshold or the temperature does not
// trigger a recompile.  Values live in registers (`v<slot>[8]`, the 8 samples a thread owns);
// windows of the same nesting depth share one barrier.
//
// Host-only C++ (also compiled into the test-only emulator library).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pystl_b200.h"
#include "fused_host.h"
#include "program.h"

namespace pystl {
namespace {

struct Src {
  std::string s;
  int indent = 1;
  void line(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    for (int i = 0; i < indent; ++i) s += "  ";
    s += buf;
    s += "\n";
  }
};

int roundup(int x, int m) { return ((x + m - 1) / m) * m; }
int clamp_off(int x, int lext) { return x > lext ? lext : (x < -lext ? -lext : x); }

// mirror of fused::WinOp<R, WA, WB, NT>::FLOATS (checked by a static_assert in the generated code)
int win_floats(int aw, int wa, int wb, int nt) {
  const int W = wb - wa + 1;
  const bool vh = W > 8;
  const int GL = wa < 0 ? roundup(-wa, 8) : 0;
  const int GR = (wb > 0 ? roundup(wb, 8) : 0) + 8;
  const int PLANE = GL + 8 * nt + GR;
  const bool use_suf = vh ? (wa != 0) : true;
  const bool use_pre = vh && (wb != 0);
  const int npl = (use_suf ? 1 : 0) + (use_pre ? 1 : 0);
  const int aggn = vh ? roundup(PLANE / 8, 4) : 0;
  return aw * (npl * PLANE + aggn);
}

const char* sem_name(int sem) {
  switch (sem) {
    case SEM_CLASSIC: return "SEM_CLASSIC";
    case SEM_SMOOTH: return "SEM_SMOOTH";
    case SEM_CUMULATIVE: return "SEM_CUMULATIVE";
    case SEM_AGM: return "SEM_AGM";
    case SEM_DGMSR: return "SEM_DGMSR";
    default: return "SEM_CLASSIC_SRC";
  }
}
const char* mode_name(int m) { return m == MODE_AND ? "MODE_AND" : (m == MODE_OR ? "MODE_OR" : "MODE_SUM"); }

bool is_window(const TInstr& in) { return in.op == T_WIN_DIRECT || in.op == T_WIN_VH; }
bool is_scan(const TInstr& in) { return in.op == T_CLAMP_SCAN; }
bool is_until(const TInstr& in) { return in.op == T_UNTIL && !in.unbounded; }
bool is_wwin(const TInstr& in) { return in.op == T_WIN_WEIGHTED; }
// instructions whose result needs one barrier (phase 1 | barrier | phase 2)
bool is_staged(const TInstr& in) { return is_window(in) || is_scan(in) || is_until(in) || is_wwin(in); }

// window formulas (no Until) under Smooth, value + gradient: latency bound, they gain from 27 instead of 24 resident warps
constexpr int kWideSmoothThreads = 288;
static bool wide_smooth_backward(const TileProgram& P) {
  if (P.sem != SEM_SMOOTH) return false;
  for (int i = 0; i < P.n_instr; ++i)
    if (P.ins[i].op == T_UNTIL) return false;
  return true;
}

struct Gen {
  const Lowered& lw;
  const TileProgram& P;
  int D, nt, minb;
  Src o;
  std::vector<int> level;     // per slot
  std::vector<int> ins_level; // per instruction: stage in which its result becomes available
  std::vector<int> scr_off;   // per instruction: scratch offset (floats)
  std::vector<int> stash_off; // per instruction: smooth operand stash
  int smem_floats = 0;
  int red_floats = 0;
  int n_stages = 0;
  bool bwd = false;                // generating the backward kernel
  int fold_root = -1;              // backward: the root n-ary whose adjoint is folded into its forward loop (no partial arrays)
  std::vector<int> bscr_off;       // per instruction: scratch of the adjoint window
  int flag_off = 0;                // smooth backward: per-window lost flags
  std::vector<int> fix_off;        // smooth backward: per window, the exact-scatter plane (aliases the dead forward scratch of its level)
  int sp_off = 0;                  // cumulative backward: four scatter planes
  std::vector<int> cum_sp;         // cumulative backward: per instruction, its borrowed scatter / hand-off planes
  std::vector<int> hand_off, aux_off;  // lockstep until adjoints: hand-off planes (dead forward scratch of the level) / staged operands (the other region)
  int stage_off = 0;               // cooperative row staging (aliases the scratch regions)
  bool coop_stage() const { return D >= 4 && D <= 8; }
  // D == 2: the kernel can take its tile through one TMA tensor copy (run-time flag TileGeom::use_tma)
  bool tma_capable() const {
    if (D != 2 && D != 4 && D != 8) return false;
    const int rows = nt * (D == 8 ? 2 : 1);  // 64-byte rows for D = 2, 128-byte rows for D = 4, 8
    return rows % ((rows + 255) / 256) == 0;
  }
  bool tma_store_capable() const { return D == 2 && tma_capable(); }
  int bar_off = 0;                 // the TMA load's mbarrier
  bool cum_bwd() const { return bwd && P.sem == SEM_CUMULATIVE; }
  // classic value-and-gradient: the forward program runs on (value, source) pairs (SEM_CLASSIC_SRC,
  // tile_eval.h) and the gradient is ONE scatter of the root cotangent through the recorded leaf
  bool src_bwd() const { return bwd && P.sem == SEM_CLASSIC; }
  bool src_image() const { return src_bwd() && D <= 4; }
  bool has_until() const {
    for (int i = 0; i < P.n_instr; ++i)
      if (is_until(P.ins[i])) return true;
    return false;
  }
  bool arg_window(const TInstr& in) const { return cum_bwd() && is_window(in) && in.mode != MODE_SUM; }
  // Until adjoints on the lockstep shared-tail walk (fused_rt.cuh: smooth_until_adj_walk, agm_until_adj_walk)
  bool lockstep_adj(const TInstr& in) const {
    if (!bwd || !is_until(in)) return false;
    if (P.sem == SEM_SMOOTH) return true;
    const bool uniform = in.w_off < 0 && in.w2_off < 0 && in.a <= 64;
    // (dgmsr: p = 1 is baked into the kernel; unit weights_pair are a run-time condition, the per-output sweep stays as the other branch)
    return uniform && (P.sem == SEM_AGM || (P.sem == SEM_DGMSR && P.p == 1));
  }
  // forward Smooth Until (unweighted, the shared-tail walk): lambda / rho exponentiated once per sample into two more planes
  bool until_expl(const TInstr& in) const {
    static const bool off = getenv("PYSTL_B200_UNTIL_EXP") && atoi(getenv("PYSTL_B200_UNTIL_EXP")) == 0;
    return !off && !bwd && P.sem == SEM_SMOOTH && is_until(in) && in.a <= 64 && in.w_off < 0 && in.w2_off < 0;
  }
  int until_floats(const TInstr& in) const { return (2 * vw() + (until_expl(in) ? 2 : 0)) * 8 * nt; }
  bool cum_lockstep(const TInstr& in) const { return cum_bwd() && is_until(in) && in.a <= 64; }
  // cumulative backward: floats of dead forward scratch an instruction's adjoint borrows (scatter / hand-off planes)
  int cum_sp_floats(const TInstr& in) const {
    if (cum_lockstep(in)) return 3 * 8 * nt + 2 * roundup(nt, 32);
    if (arg_window(in)) return 2 * 8 * nt;
    return 0;
  }
  int hand_planes() const { return P.sem == SEM_SMOOTH ? 2 : 3; }
  int aux_floats() const { return P.sem == SEM_SMOOTH ? 2 * 8 * nt + roundup(nt, 32) : 2 * roundup(nt, 32); }

  Gen(const Lowered& l, int D_, int nt_, int minb_, bool bwd_) : lw(l), P(l.prog), D(D_), nt(nt_), minb(minb_), bwd(bwd_) {}
  int kw() const { return P.sem == SEM_SMOOTH ? 1 : 2; }  // words of an adjoint coefficient
  // Smooth value-and-gradient kernels of window formulas on the wide tile (fused_geometry: 288 threads, 3 CTAs per SM at <= 75
  // registers): the lifted operands are not kept in 8 registers per window for the adjoint, they are lifted again from the stash
  // -- that is what lets 27 warps per SM run at 72 registers with 16 bytes of stack (cfg2: 1.602 -> 1.555 ms, as the default 1.560 ms).  PYSTL_B200_RELIFT = 0 / 1 overrides.
  bool relift() const {
    const char* e = getenv("PYSTL_B200_RELIFT");
    if (e && *e) return atoi(e) != 0;
    return bwd && nt == kWideSmoothThreads && wide_smooth_backward(P);
  }

  int vw() const { return (P.sem == SEM_CUMULATIVE || src_bwd()) ? 2 : 1; }
  int aw() const { return src_bwd() ? 2 : ((P.sem == SEM_CLASSIC || P.sem == SEM_SMOOTH) ? 1 : 2); }

  bool supported(std::string* why) const {
    if (D < 1) {
      *why = "bad D";
      return false;
    }
    for (int i = 0; i < P.n_instr; ++i) {
      const TInstr& in = P.ins[i];
      if (is_scan(in) && P.sem == SEM_CLASSIC) continue;
      if (is_until(in) && (!bwd || P.sem == SEM_CUMULATIVE || in.b <= 1024)) continue;
      if (is_wwin(in)) continue;
      if (in.op != T_NARY && !is_window(in)) {
        *why = "until / open-ended windows are evaluated by the interpreter tier";
        return false;
      }
    }
    return true;
  }

  void plan() {
    level.assign(kMaxSlots, 0);
    ins_level.assign(P.n_instr, 0);
    scr_off.assign(P.n_instr, 0);
    stash_off.assign(P.n_instr, -1);
    int max_level = 0;
    for (int i = 0; i < P.n_instr; ++i) {
      const TInstr& in = P.ins[i];
      int l = 0;
      for (int j = 0; j < in.nsrc; ++j) l = std::max(l, level[in.src[j]]);
      if (is_staged(in)) l += 1;
      level[in.dst] = l;
      ins_level[i] = l;
      max_level = std::max(max_level, l);
    }
    n_stages = max_level + 1;
    // shared memory: [reduction scratch][persistent][region 0][region 1].  Forward windows of level
    // l use region l & 1 (the barrier of level l+1 separates the readers of level l from the
    // writers of level l+2); the adjoint windows of level l use region (l+1) & 1 for the same
    // reason, walking the levels downwards.
    const int nw = nt / 32;
    red_floats = roundup(std::max(2 * std::max(P.n_preds, 1), 1) * nw, 32);
    if (tma_capable()) {
      bar_off = red_floats;
      red_floats += 32;
    }
    red_floats += 32;  // + the mean reduction's own 32 floats
    bscr_off.assign(P.n_instr, 0);
    int persist = 0;
    const bool stash_persists = bwd && P.sem == SEM_SMOOTH;  // the exact scatter re-reads the operands
    if (stash_persists) {
      for (int i = 0; i < P.n_instr; ++i)
        if (is_window(P.ins[i])) {
          stash_off[i] = red_floats + persist;
          persist += 8 * nt;
        }
      flag_off = red_floats + persist;
      persist += roundup(P.n_instr, 4);
    }
    if (src_bwd()) {
      persist = roundup(red_floats + persist, 32) - red_floats;  // (128-byte aligned for the TMA store)
      sp_off = red_floats + persist;
      // narrow signals: the gradient tile as it lies in HBM (8 * nt rows of D floats); wide ones: one
      // scatter plane per predicate (a 64 KB image would halve the resident CTAs on cfg5)
      persist += src_image() ? 8 * nt * D : P.n_preds * 8 * nt;
    }
    if (cum_bwd() || (bwd && has_until())) {
      // persistent scatter planes only for the per-output until adjoints; the lockstep walks and the cumulative
      // arg-min / arg-max scatters take the dead forward scratch of their level (cum_sp below)
      bool need_sp = false;
      for (int i = 0; i < P.n_instr; ++i) need_sp |= is_until(P.ins[i]) && !lockstep_adj(P.ins[i]) && !cum_lockstep(P.ins[i]);
      if (need_sp) {
        sp_off = red_floats + persist;
        persist += 4 * 8 * nt;
      }
      for (int i = 0; i < P.n_instr; ++i)
        if (is_until(P.ins[i])) {  // the adjoint re-reads the staged operands
          scr_off[i] = red_floats + persist;
          persist += 2 * vw() * 8 * nt;
        }
    }
    std::vector<int> used(max_level + 3, 0), bused(max_level + 3, 0);
    for (int i = 0; i < P.n_instr; ++i) {
      const TInstr& in = P.ins[i];
      if (is_scan(in)) {
        scr_off[i] = used[ins_level[i]];
        used[ins_level[i]] += (src_bwd() ? 4 : 2) * 32;  // ClampScanOp<SEM>::FLOATS
        continue;
      }
      if (is_wwin(in)) {
        scr_off[i] = used[ins_level[i]];
        used[ins_level[i]] += 2 * 8 * nt;  // WeightedWinOp::FLOATS
        if (bwd) {
          bscr_off[i] = bused[ins_level[i]];
          bused[ins_level[i]] += 2 * 8 * nt;  // WeightedWinAdj::FLOATS
        }
        continue;
      }
      if (is_until(in)) {
        if (bwd) continue;  // persistent (above): the adjoint re-reads the staged operands
        scr_off[i] = used[ins_level[i]];
        used[ins_level[i]] += until_floats(in);  // UntilOp::FLOATS
        continue;
      }
      if (!is_window(in)) continue;
      const int l = ins_level[i];
      const int a = clamp_off(in.a, 8 * nt), b = clamp_off(in.b, 8 * nt);
      scr_off[i] = used[l];
      used[l] += arg_window(in) ? 2 * win_floats(2, a, b, nt) : win_floats(aw(), a, b, nt);
      if (P.sem == SEM_SMOOTH && !stash_persists) {
        stash_off[i] = used[l];
        used[l] += 8 * nt;
      }
      if (bwd && !arg_window(in) && !src_bwd()) {
        bscr_off[i] = bused[l];
        bused[l] += win_floats(kw(), -b, -a, nt);
      }
    }
    int reg[2] = {0, 0};
    for (int l = 0; l <= max_level; ++l) {
      reg[l & 1] = std::max(reg[l & 1], used[l]);
      reg[(l + 1) & 1] = std::max(reg[(l + 1) & 1], bused[l]);
    }
    // smooth backward: the exact scatter of a window of level l (rare path of its adjoint) needs one plane; the
    // forward scratch of that level (region l & 1) is dead by then and the adjoint windows live in the other region
    fix_off.assign(P.n_instr, 0);
    if (stash_persists)
      for (int i = 0; i < P.n_instr; ++i)
        if (is_window(P.ins[i])) reg[ins_level[i] & 1] = std::max(reg[ins_level[i] & 1], 8 * nt);
    // lockstep until adjoints of level l: the hand-off planes take the dead forward scratch of that level (region
    // l & 1); the staged operands take the other region, which the adjoint windows of level l only fill afterwards
    hand_off.assign(P.n_instr, 0);
    aux_off.assign(P.n_instr, 0);
    for (int i = 0; i < P.n_instr; ++i)
      if (lockstep_adj(P.ins[i])) {
        const int l = ins_level[i];
        reg[l & 1] = std::max(reg[l & 1], hand_planes() * 8 * nt);
        reg[(l + 1) & 1] = std::max(reg[(l + 1) & 1], aux_floats());
      }
    cum_sp.assign(P.n_instr, 0);
    if (cum_bwd())
      for (int i = 0; i < P.n_instr; ++i) reg[ins_level[i] & 1] = std::max(reg[ins_level[i] & 1], cum_sp_floats(P.ins[i]));
    const int base0 = red_floats + persist, base1 = base0 + reg[0];
    for (int i = 0; i < P.n_instr; ++i) {
      cum_sp[i] = (ins_level[i] & 1) ? base1 : base0;
      if (is_window(P.ins[i])) fix_off[i] = (ins_level[i] & 1) ? base1 : base0;
      if (lockstep_adj(P.ins[i])) {
        hand_off[i] = (ins_level[i] & 1) ? base1 : base0;
        aux_off[i] = ((ins_level[i] + 1) & 1) ? base1 : base0;
      }
    }
    for (int i = 0; i < P.n_instr; ++i) {
      if (is_scan(P.ins[i]) || is_wwin(P.ins[i]) || (is_until(P.ins[i]) && !bwd)) scr_off[i] += (ins_level[i] & 1) ? base1 : base0;
      if (is_wwin(P.ins[i])) bscr_off[i] += ((ins_level[i] + 1) & 1) ? base1 : base0;
      if (!is_window(P.ins[i])) continue;
      const int l = ins_level[i];
      scr_off[i] += (l & 1) ? base1 : base0;
      if (stash_off[i] >= 0 && !stash_persists) stash_off[i] += (l & 1) ? base1 : base0;
      bscr_off[i] += ((l + 1) & 1) ? base1 : base0;
    }
    smem_floats = base1 + reg[1];
    // staging area of the signal rows (cooperative loads, TMA box; the TMA store of the gradient rows at the very
    // end): nothing else is alive in shared memory at either time except the reduction scratch, so it starts right
    // behind that (1024-byte aligned for the swizzled TMA box)
    if (coop_stage() || tma_capable()) {
      stage_off = roundup(red_floats, 256);
      smem_floats = std::max(smem_floats, stage_off + 8 * nt * D);
    }
  }

  // expression of operand j of instruction `in` at sample i
  std::string operand(const TInstr& in, int j, const char* idx) const {
    char buf[96];
    const bool neg = (in.negmask >> j) & 1;
    if (neg) snprintf(buf, sizeof(buf), "negate<SEM>(v%d[%s])", (int)in.src[j], idx);
    else snprintf(buf, sizeof(buf), "v%d[%s]", (int)in.src[j], idx);
    return buf;
  }

  // ---- which values can be NaN: an Until without a candidate (lower bound above the rest of the trace) and a weighted
  // window whose clipped weights sum to zero write NaN where the reference raises, and NaN flows through everything above
  std::vector<char> may_nan;
  void plan_nan() {
    may_nan.assign(kMaxSlots, 0);
    for (int i = 0; i < P.n_instr; ++i) {
      const TInstr& in = P.ins[i];
      bool n = (is_until(in) && in.a > 0) || is_wwin(in);
      for (int j = 0; j < in.nsrc; ++j) n |= may_nan[in.src[j]] != 0;
      may_nan[in.dst] = n;
    }
  }

  // ---- smooth only: which value-range bounds are consumed (a window shifts its exponentials by
  // a bound of its operand; bounds propagate analytically through the operators above)
  std::vector<char> need_lo, need_hi;
  void plan_bounds() {
    need_lo.assign(kMaxSlots, 0);
    need_hi.assign(kMaxSlots, 0);
    if (P.sem != SEM_SMOOTH) return;
    for (int i = P.n_instr - 1; i >= 0; --i) {
      const TInstr& in = P.ins[i];
      const bool want_lo = need_lo[in.dst], want_hi = need_hi[in.dst];
      for (int j = 0; j < in.nsrc; ++j) {
        const bool neg = (in.negmask >> j) & 1;
        bool olo = want_lo, ohi = want_hi;  // bounds of the (negated) operand that are consumed
        if (is_window(in)) (in.mode == MODE_AND ? olo : ohi) = true;  // the shift
        if (is_until(in)) {  // shift by the smaller operand; the result's range mixes all four bounds
          olo = true;
          ohi = want_lo || want_hi;
          if (ohi) olo = true;
        }
        if (olo) (neg ? need_hi : need_lo)[in.src[j]] = 1;
        if (ohi) (neg ? need_lo : need_hi)[in.src[j]] = 1;
      }
    }
  }
  std::string lo_of(int slot, bool neg) const { return (neg ? "-hi" : "lo") + std::to_string(slot); }
  std::string hi_of(int slot, bool neg) const { return (neg ? "-lo" : "hi") + std::to_string(slot); }

  void emit_predicates() {
    o.line("// ---- predicates: rows -> registers (api.py:142-143)");
    for (int q = 0; q < P.n_preds; ++q) o.line("V v%d[8];", (int)P.preds[q].dst);
    if (tma_capable()) {
      o.line("float r[8][D];");
      o.line("if (g.use_tma) load_rows_tma<D, NT>(c, g, smem + %d, smem + %d, r);  // one TMA tensor copy per CTA", stage_off, bar_off);
    }
    o.line("if (live) {");
    o.indent++;
    if (D <= 8) {
      if (!tma_capable()) o.line("float r[8][D];");
      if (coop_stage())
        o.line("%sload_rows_coop<D>(c, signal, g.T, smem + %d, r);  // coalesced, transposed through shared memory",
               tma_capable() ? "if (!g.use_tma) " : "", stage_off);
      else if (tma_capable()) o.line("if (!g.use_tma) load_rows<D>(c, signal, g.T, r);");
      else o.line("load_rows<D>(c, signal, g.T, r);");
      o.line("prefetch_rows<D>(c, signal, g.B, g.T, g.scr_planes);  // (scr_planes carries the prefetch distance in traces)");
    }
    for (int q = 0; q < P.n_preds; ++q) {
      const TPred& pr = P.preds[q];
      if (D > 8) {
        for (int k = 0; k < pr.n_terms; ++k) {
          o.line("float x%d_%d[8];", q, k);
          o.line("load_channel(c, signal, g.T, D, %d, x%d_%d);", (int)pr.ch[k], q, k);
        }
      }
      o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i) {");
      o.indent++;
      auto term = [&](int k) {
        char b[64];
        if (D <= 8) snprintf(b, sizeof(b), "r[i][%d]", (int)pr.ch[k]);
        else snprintf(b, sizeof(b), "x%d_%d[i]", q, k);
        return std::string(b);
      };
      const bool f64 = P.sem == SEM_CLASSIC && pr.f64;
      if (pr.pair) {
        o.line("v%d[i] = V{{%s, %s}};  // a cumulative value computed by the host: (pos, neg) as they stand", (int)pr.dst, term(0).c_str(),
               term(1).c_str());
        o.indent--;
        o.line("}");
        continue;
      }
      if (f64) {
        // same operations and order as the reference's float64 evaluation, one final rounding
        o.line("double a = 0.0;");
        for (int k = 0; k < pr.n_terms; ++k)
          o.line("a = fma((double)P.preds[%d].coeff[%d], (double)%s, a);", q, k, term(k).c_str());
        o.line("const float acc = (float)(a + (double)P.preds[%d].bias);", q);
      } else {
        // (coeff * x + bias in ONE rounding for a single term: what "+-x + c" needs to match the reference bit for bit)
        if (pr.n_terms == 1) {
          o.line("const float acc = fmaf(P.preds[%d].coeff[0], %s, P.preds[%d].bias);", q, term(0).c_str(), q);
        } else {
          o.line("float acc = 0.f;");
          for (int k = 0; k < pr.n_terms; ++k) o.line("acc = fmaf(P.preds[%d].coeff[%d], %s, acc);", q, k, term(k).c_str());
          o.line("acc += P.preds[%d].bias;", q);
        }
      }
      if (P.sem == SEM_CUMULATIVE) o.line("v%d[i] = V{{fmaxf(0.f, acc), fminf(0.f, acc)}};  // cumulative.py:29-30", (int)pr.dst);
      else if (src_bwd()) o.line("v%d[i] = V{{acc, i2f(%d | (c.p0 + i))}};  // source: predicate << 24 | tile position", (int)pr.dst, q << 24);
      else o.line("v%d[i] = V{{acc}};", (int)pr.dst);
      o.indent--;
      o.line("}");
    }
    o.indent--;
    o.line("}");
    if (coop_stage()) o.line("simt::sync();  // the staging area is reused as window scratch");
    else if (tma_capable()) o.line("if (g.use_tma) simt::sync();  // the staging area is reused as window scratch");
    if (P.sem == SEM_SMOOTH) {
      int nb = 0;
      for (int q = 0; q < P.n_preds; ++q) nb += need_lo[P.preds[q].dst] + need_hi[P.preds[q].dst];
      if (nb > 0) {
        o.line("// value range of the predicates over the tile: the windows shift their exponentials by it");
        o.line("float bm[%d];", nb);
        o.line("_Pragma(\"unroll\") for (int q = 0; q < %d; ++q) bm[q] = -f_inf();", nb);
        o.line("if (live) {");
        o.indent++;
        int k = 0;
        for (int q = 0; q < P.n_preds; ++q) {
          const int d = P.preds[q].dst;
          if (need_lo[d]) o.line("bm[%d] = max8<true>(c, v%d);", k++, d);
          if (need_hi[d]) o.line("bm[%d] = max8<false>(c, v%d);", k++, d);
        }
        o.indent--;
        o.line("}");
        o.line("block_max<NT, %d>(c, red, bm);", nb);
        k = 0;
        for (int q = 0; q < P.n_preds; ++q) {
          const int d = P.preds[q].dst;
          if (need_lo[d]) o.line("const float lo%d = -bm[%d];", d, k++);
          if (need_hi[d]) o.line("const float hi%d = bm[%d];", d, k++);
        }
      }
    }
  }

  void win_decl(int i) {
    const TInstr& in = P.ins[i];
    const int a = clamp_off(in.a, 8 * nt), b = clamp_off(in.b, 8 * nt);
    o.line("using R%d = Red<SEM, %s>;", i, mode_name(in.mode));
    o.line("using W%d = WinOp<R%d, %d, %d, NT>;", i, i, a, b);
    o.line("static_assert(W%d::FLOATS == %d, \"host / device scratch layout mismatch\");", i, win_floats(aw(), a, b, nt));
  }

  // cumulative backward: minima / maxima on (value, position) pairs, one channel at a time
  void argwin_phase1(int i) {
    const TInstr& in = P.ins[i];
    const int a = clamp_off(in.a, 8 * nt), b = clamp_off(in.b, 8 * nt);
    const int half = win_floats(2, a, b, nt);
    o.line("// window %d (arg%s), phase 1", i, in.mode == MODE_AND ? "min" : "max");
    o.line("using W%d = WinOp<ArgRed<%s>, %d, %d, NT>;", i, mode_name(in.mode), a, b);
    o.line("static_assert(W%d::FLOATS == %d, \"host / device scratch layout mismatch\");", i, half);
    o.line("Vec<2> xa%d[8], sfa%d[8], pfa%d[8], xb%d[8], sfb%d[8], pfb%d[8];", i, i, i, i, i, i);
    o.line("W%d::fill_guards(smem + %d, c.tid);", i, scr_off[i]);
    o.line("W%d::fill_guards(smem + %d, c.tid);", i, scr_off[i] + half);
    o.line("if (live) {");
    o.indent++;
    o.line("V t[8];");
    o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i) t[i] = %s;", operand(in, 0, "i").c_str());
    o.line("arg_lift8<%s>(c, t, 0, xa%d);", mode_name(in.mode), i);
    o.line("arg_lift8<%s>(c, t, 1, xb%d);", mode_name(in.mode), i);
    o.line("W%d::phase1(smem + %d, c.tid, xa%d, sfa%d, pfa%d);", i, scr_off[i], i, i, i);
    o.line("W%d::phase1(smem + %d, c.tid, xb%d, sfb%d, pfb%d);", i, scr_off[i] + half, i, i, i);
    o.indent--;
    o.line("} else {");
    o.indent++;
    o.line("W%d::phase1_dead(smem + %d, c.tid);", i, scr_off[i]);
    o.line("W%d::phase1_dead(smem + %d, c.tid);", i, scr_off[i] + half);
    o.indent--;
    o.line("}");
  }
  void argwin_phase2(int i) {
    const TInstr& in = P.ins[i];
    const int a = clamp_off(in.a, 8 * nt), b = clamp_off(in.b, 8 * nt);
    const int half = win_floats(2, a, b, nt);
    const int d = in.dst;
    o.line("// window %d (arg), phase 2", i);
    o.line("V v%d[8];", d);
    o.line("Vec<2> acca%d[8], accb%d[8];  // (value, position) of both channels, kept for the adjoint", i, i);
    o.line("if (live) {");
    o.indent++;
    o.line("W%d::phase2(smem + %d, c.tid, xa%d, sfa%d, pfa%d, acca%d);", i, scr_off[i], i, i, i, i);
    o.line("W%d::phase2(smem + %d, c.tid, xb%d, sfb%d, pfb%d, accb%d);", i, scr_off[i] + half, i, i, i, i);
    o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i) v%d[i] = V{{acca%d[i].x[0], accb%d[i].x[0]}};", d, i, i);
    o.indent--;
    o.line("}");
  }

  void win_phase1(int i) {
    const TInstr& in = P.ins[i];
    if (arg_window(in)) return argwin_phase1(i);
    const bool neg = in.negmask & 1;
    const int s = in.src[0];
    o.line("// window %d, phase 1: lift, chunk scans -> shared", i);
    win_decl(i);
    o.line("R%d::A x%d[8], sf%d[8], pf%d[8];", i, i, i, i);
    o.line("RP rp%d = rp;", i);
    if (P.sem == SEM_SMOOTH) {
      if (in.mode == MODE_AND) o.line("rp%d.shift = -(%s) * rp.k;  // max exponent of the operand over the tile", i, lo_of(s, neg).c_str());
      else o.line("rp%d.shift = (%s) * rp.k;", i, hi_of(s, neg).c_str());
    }
    o.line("W%d::fill_guards(smem + %d, c.tid);", i, scr_off[i]);
    o.line("if (live) {");
    o.indent++;
    o.line("V t[8];");
    o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i) t[i] = %s;", operand(in, 0, "i").c_str());
    o.line("lift8<R%d, VW>(c, rp%d, t, x%d);", i, i, i);
    if (P.sem == SEM_SMOOTH) o.line("stash8<VW>(smem + %d, c.tid, t, 0);", stash_off[i]);
    o.line("W%d::phase1(smem + %d, c.tid, x%d, sf%d, pf%d);", i, scr_off[i], i, i, i);
    o.indent--;
    o.line("} else {");
    o.indent++;
    o.line("W%d::phase1_dead(smem + %d, c.tid);", i, scr_off[i]);
    o.indent--;
    o.line("}");
  }

  void win_phase2(int i) {
    const TInstr& in = P.ins[i];
    if (arg_window(in)) return argwin_phase2(i);
    const int a = clamp_off(in.a, 8 * nt), b = clamp_off(in.b, 8 * nt);
    const bool neg = in.negmask & 1;
    const int s = in.src[0], d = in.dst;
    o.line("// window %d, phase 2: combine -> value", i);
    o.line("V v%d[8];", d);
    if (bwd) o.line("R%d::A acc%d[8];  // kept for the adjoint", i, i);
    o.line("if (live) {");
    o.indent++;
    if (bwd) o.line("R%d::A (&acc)[8] = acc%d;", i, i);
    else o.line("R%d::A acc[8];", i);
    o.line("W%d::phase2(smem + %d, c.tid, x%d, sf%d, pf%d, acc);", i, scr_off[i], i, i, i);
    o.line("fin8<R%d, %d, %d, VW>(c, rp%d, acc, v%d);", i, a, b, i, d);
    if (P.sem == SEM_SMOOTH)
      o.line("smooth_fix8<%s, %d, %d, NT>(c, rp%d, smem + %d, acc, v%d);", mode_name(in.mode), a, b, i, stash_off[i], d);
    o.indent--;
    o.line("}");
    if (P.sem == SEM_SMOOTH) {
      if (need_lo[d]) o.line("const float lo%d = smooth_lo<%s>(rp, %s, %d.f);", d, mode_name(in.mode), lo_of(s, neg).c_str(), b - a + 1);
      if (need_hi[d]) o.line("const float hi%d = smooth_hi<%s>(rp, %s, %d.f);", d, mode_name(in.mode), hi_of(s, neg).c_str(), b - a + 1);
    }
  }

  void scan_phase1(int i) {
    const TInstr& in = P.ins[i];
    o.line("// classic until %d: clamp scan, phase 1 (chunk composition, warp scan)", i);
    o.line("using CS%d = ClampScanOp<SEM, NT>;", i);
    o.line("static_assert(CS%d::FLOATS == %d, \"host / device scratch layout mismatch\");", i, src_bwd() ? 128 : 64);
    o.line("CS%d::A cx%d[8], ce%d;", i, i, i);
    o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i)");
    o.line("  cx%d[i] = (live && c.p0 + i < c.nlive) ? ClampOp<SEM>::make(%s, %s) : ClampOp<SEM>::ident();", i, operand(in, 0, "i").c_str(),
           operand(in, 1, "i").c_str());
    o.line("CS%d::phase1(smem + %d, c, cx%d, ce%d);", i, scr_off[i], i, i);
  }
  void scan_phase2(int i) {
    const TInstr& in = P.ins[i];
    o.line("// classic until %d: clamp scan, phase 2", i);
    o.line("V v%d[8];", (int)in.dst);
    o.line("CS%d::phase2<VW>(smem + %d, c, cx%d, ce%d, v%d);", i, scr_off[i], i, i, (int)in.dst);
  }

  void wwin_phase1(int i) {
    const TInstr& in = P.ins[i];
    o.line("// weighted window %d [%d, %d], phase 1: stage the unweighted lifted terms", i, in.a, in.b);
    o.line("using WW%d = WeightedWinOp<SEM, %s, %d, %d, NT>;", i, mode_name(in.mode), in.a, in.b);
    o.line("static_assert(WW%d::FLOATS == %d, \"host / device scratch layout mismatch\");", i, 2 * 8 * nt);
    o.line("if (live) {");
    o.indent++;
    o.line("V t[8];");
    o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i) t[i] = %s;", operand(in, 0, "i").c_str());
    o.line("WW%d::A x[8];", i);
    o.line("lift8<WW%d::R, VW>(c, rp, t, x);", i);
    o.line("WW%d::phase1(smem + %d, c.tid, x);", i, scr_off[i]);
    o.indent--;
    o.line("}");
  }
  void wwin_phase2(int i) {
    const TInstr& in = P.ins[i];
    o.line("// weighted window %d, phase 2: every output folds its window with its own weights", i);
    o.line("V v%d[8];", (int)in.dst);
    o.line("if (live) WW%d::phase2(smem + %d, c, rp, P.wbuf + %d, v%d);", i, scr_off[i], (int)in.w_off, (int)in.dst);
  }

  void until_phase1(int i) {
    const TInstr& in = P.ins[i];
    const bool nl = in.negmask & 1, nr = in.negmask & 2;
    o.line("// until %d [%d, %d], phase 1: stage both operands", i, in.a, in.b);
    o.line("using U%d = UntilOp<SEM, %d, %d, NT%s>;", i, in.a, in.b, until_expl(in) ? ", true" : "");
    o.line("static_assert(U%d::FLOATS == %d, \"host / device scratch layout mismatch\");", i, bwd ? 2 * vw() * 8 * nt : until_floats(in));
    o.line("RP rp%d = rp;", i);
    if (P.sem == SEM_SMOOTH)
      o.line("rp%d.shift = fmaxf(-(%s), -(%s)) * rp.k;  // max exponent of -l, -r over the tile", i, lo_of(in.src[0], nl).c_str(),
             lo_of(in.src[1], nr).c_str());
    o.line("if (live) {");
    o.indent++;
    o.line("V tl[8], tr[8];");
    o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i) { tl[i] = %s; tr[i] = %s; }", operand(in, 0, "i").c_str(), operand(in, 1, "i").c_str());
    if (until_expl(in)) o.line("U%d::phase1(smem + %d, c.tid, tl, tr, rp%d);", i, scr_off[i], i);
    else o.line("U%d::phase1(smem + %d, c.tid, tl, tr);", i, scr_off[i]);
    o.indent--;
    o.line("}");
  }
  void until_phase2(int i) {
    const TInstr& in = P.ins[i];
    const int d = in.dst;
    const bool nl = in.negmask & 1, nr = in.negmask & 2;
    o.line("// until %d, phase 2: O(b) walk with running state", i);
    o.line("V v%d[8];", d);
    const bool keep_a = bwd && P.sem == SEM_SMOOTH;
    if (keep_a) o.line("float ua%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // A_t = sum of the candidates' terms, kept for the adjoint", i);
    o.line("if (live) {");
    o.indent++;
    if (in.w_off >= 0 || in.w2_off >= 0) {
      char wl[48] = "nullptr", wr[48] = "nullptr";
      if (in.w_off >= 0) snprintf(wl, sizeof(wl), "P.wbuf + %d", (int)in.w_off);
      if (in.w2_off >= 0) snprintf(wr, sizeof(wr), "P.wbuf + %d", (int)in.w2_off);
      o.line("until_weighted_phase2<SEM, %d, %d, NT>(smem + %d, c, rp%d, P.ins[%d].w[0], P.ins[%d].w[1], %s, %s, v%d);", in.a, in.b,
             scr_off[i], i, i, i, wl, wr, d);
    } else {
      o.line("UntilState<SEM> st[8];");
      o.line("U%d::phase2(smem + %d, c, rp%d, P.ins[%d].w[0], P.ins[%d].w[1], v%d, st);", i, scr_off[i], i, i, i, d);
      if (P.sem == SEM_SMOOTH) o.line("U%d::smooth_fix(smem + %d, c, rp%d, st, v%d);", i, scr_off[i], i, d);
      if (keep_a) o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i) ua%d[i] = st[i].acc;", i);
    }
    o.indent--;
    o.line("}");
    if (P.sem == SEM_SMOOTH && (need_lo[d] || need_hi[d])) {
      // softmin over <= b+2 terms inside a softmax over <= b+1 candidates (tile_eval.h: op_until)
      o.line("const float pad%d = 2.f * rp.invk * fast_lg2(%d.f) + 2e-6f * (fabsf(%s) + fabsf(%s)) + 1e-30f;", i, in.b + 2,
             lo_of(in.src[0], nl).c_str(), hi_of(in.src[1], nr).c_str());
      if (need_lo[d]) o.line("const float lo%d = fminf(%s, %s) - pad%d;", d, lo_of(in.src[0], nl).c_str(), lo_of(in.src[1], nr).c_str(), i);
      if (need_hi[d]) o.line("const float hi%d = fmaxf(%s, %s) + pad%d;", d, hi_of(in.src[0], nl).c_str(), hi_of(in.src[1], nr).c_str(), i);
    }
  }

  void nary(int i) {
    const TInstr& in = P.ins[i];
    const int n = in.nsrc, d = in.dst;
    o.line("// n-ary %s (%d operands)", in.mode == MODE_AND ? "And" : "Or", n);
    o.line("V v%d[8];", d);
    const bool fold = bwd && fold_root == i;  // the cotangent of the root is known now: multiply at once
    if (bwd && !cum_bwd() && !src_bwd() && !fold)
      for (int j = 0; j < n; ++j) o.line("float d%d_%d[8];  // d out / d operand %d, for the adjoint sweep", i, j, j);
    o.line("if (live) {");
    o.indent++;
    if (fold) o.line("root_cotangent<%s>(c, g.T, grad_out, scale, g%d);", P.root_neg ? "true" : "false", d);
    std::string w = "const float w[" + std::to_string(n) + "] = {";
    for (int j = 0; j < n; ++j) {
      char b[64];
      snprintf(b, sizeof(b), "%sP.ins[%d].w[%d]", j ? ", " : "", i, j);
      w += b;
    }
    o.line("%s};", w.c_str());
    o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i) {");
    o.indent++;
    std::string a = "const V a[" + std::to_string(n) + "] = {";
    for (int j = 0; j < n; ++j) a += (j ? ", " : "") + operand(in, j, "i");
    o.line("%s};", a.c_str());
    if (bwd && !cum_bwd() && !src_bwd()) {
      o.line("float dd[%d];", n);
      o.line("v%d[i] = nary1_partials<SEM, %s, %d>(rp, a, w, P.ins[%d].sw, dd);", d, mode_name(in.mode), n, i);
      for (int j = 0; j < n; ++j) {
        if (fold) o.line("%s", adj_mul(in, j, d, ("dd[" + std::to_string(j) + "]").c_str()).c_str());
        else o.line("d%d_%d[i] = dd[%d];", i, j, j);
      }
    } else {
      o.line("v%d[i] = nary1<SEM, %s, %d>(rp, a, w, P.ins[%d].sw);", d, mode_name(in.mode), n, i);
    }
    o.indent--;
    o.line("}");
    o.indent--;
    o.line("}");
    if (P.sem == SEM_SMOOTH && (need_lo[d] || need_hi[d])) {
      // min / max over the operands' ranges (tile_eval.h: op_nary)
      const char* f = in.mode == MODE_AND ? "fminf" : "fmaxf";
      std::string lo, hi;
      for (int j = 0; j < n; ++j) {
        const bool neg = (in.negmask >> j) & 1;
        const std::string l = lo_of(in.src[j], neg), h = hi_of(in.src[j], neg);
        lo = j ? std::string(f) + "(" + lo + ", " + l + ")" : l;
        hi = j ? std::string(f) + "(" + hi + ", " + h + ")" : h;
      }
      if (need_lo[d]) o.line("const float lo%d = smooth_lo<%s>(rp, %s, %d.f);", d, mode_name(in.mode), lo.c_str(), n);
      if (need_hi[d]) o.line("const float hi%d = smooth_hi<%s>(rp, %s, %d.f);", d, mode_name(in.mode), hi.c_str(), n);
    }
  }

  void emit_forward_body() {
    o.line("PF_DYN_SMEM(smem);");
    o.line("const TC c = tile_begin<NT>(P, g, smem);");
    o.line("const RP rp = c.rp;");
    o.line("float* red = smem;");
    o.line("const bool live = alive(c);  // threads past the horizon only keep the shared planes consistent");
    o.line("(void)red; (void)rp;");
    emit_predicates();
    // (after the tile load: the staging area aliases them; the first barrier of the forward sweep orders the store)
    if (bwd && P.sem == SEM_SMOOTH) o.line("if (c.tid < %d) reinterpret_cast<int*>(smem + %d)[c.tid] = 0;  // lost flags", P.n_instr, flag_off);
    for (int st = 0; st < n_stages; ++st) {
      o.line("// ================= stage %d", st);
      for (int i = 0; i < P.n_instr; ++i) {
        const TInstr& in = P.ins[i];
        if (ins_level[i] != st) continue;
        if (is_window(in)) win_phase2(i);
        else if (is_scan(in)) scan_phase2(i);
        else if (is_until(in)) until_phase2(i);
        else if (is_wwin(in)) wwin_phase2(i);
        else nary(i);
      }
      bool any = false;
      for (int i = 0; i < P.n_instr; ++i)
        if (is_staged(P.ins[i]) && ins_level[i] == st + 1) {
          if (is_window(P.ins[i])) win_phase1(i);
          else if (is_until(P.ins[i])) until_phase1(i);
          else if (is_wwin(P.ins[i])) wwin_phase1(i);
          else scan_phase1(i);
          any = true;
        }
      if (any) o.line("simt::sync();");
    }
  }

  // shows up in profiler launch lists: pystl_fused_<pass>_<semantics>
  std::string kernel_name() const {
    static const char* sems[] = {"classical", "smooth", "cumulative", "agm", "dgmsr"};
    return std::string("pystl_fused_") + (bwd ? "value_and_grad_" : "forward_") + sems[P.sem];
  }

  void emit_header(const char* what) {
    o.indent = 0;
    o.line("// generated by pystl_b200/csrc/codegen.cpp -- %s kernel of one formula (fused tier)", what);
    if (P.sem == SEM_DGMSR) o.line("#define PYSTL_CONST_P %d  // dgmsr exponent, baked in", (int)P.rp.p);
    o.line("#include \"fused_rt.cuh\"");
    o.line("using namespace pystl;");
    o.line("using namespace pystl::fused;");
    o.line("constexpr int NT = %d, D = %d, SEM = %s, VW = Tr<SEM>::VW;", nt, D, src_bwd() ? "SEM_CLASSIC_SRC" : sem_name(P.sem));
    o.line("typedef Vec<VW> V;");
  }

  std::string forward() {
    plan();
    plan_bounds();
    emit_header("forward");
    o.line("PF_KERNEL(NT, %d) %s(const PF_GRID_CONSTANT TileProgram P, const PF_GRID_CONSTANT TileGeom g,", minb, kernel_name().c_str());
    o.line("    const float* __restrict__ signal, float* __restrict__ out, double* __restrict__ mean_acc) {");
    o.indent = 1;
    emit_forward_body();
    o.line("float part = 0.f;");
    o.line("if (live) part = store_root<SEM, %s>(c, g.T, v%d, out);", P.root_neg ? "true" : "false", (int)P.root);
    o.line("if (mean_acc) block_sum_to<NT>(c, part, mean_acc, smem + %d);", red_floats - 32);
    o.indent = 0;
    o.line("}");
    return o.s;
  }

  // ---------------------------------------------------------------- backward
  // g_src[i] (+)= g_dst[i] * partial.  Outputs the reference raises on are NaN with a zero cotangent: the select keeps
  // 0 * NaN out of the adjoints -- except for the two-operand smooth And / Or, whose partials are finite by construction
  std::string adj_mul(const TInstr& in, int j, int d, const char* partial) const {
    char b[160];
    const char* op = ((in.negmask >> j) & 1) ? "-=" : "+=";
    if (P.sem == SEM_SMOOTH && in.nsrc == 2) snprintf(b, sizeof(b), "g%d[i] %s g%d[i] * %s;", (int)in.src[j], op, d, partial);
    else snprintf(b, sizeof(b), "g%d[i] %s (g%d[i] != 0.f) ? g%d[i] * %s : 0.f;", (int)in.src[j], op, d, d, partial);
    return b;
  }
  void nary_bwd(int i) {
    if (i == fold_root) return;  // done in the forward sweep
    const TInstr& in = P.ins[i];
    const int n = in.nsrc, d = in.dst;
    o.line("// adjoint of the n-ary %s %d: the partials were computed with the value", in.mode == MODE_AND ? "And" : "Or", i);
    o.line("if (live) {");
    o.indent++;
    o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i) {");
    o.indent++;
    for (int j = 0; j < n; ++j) {
      char e[32];
      snprintf(e, sizeof(e), "d%d_%d[i]", i, j);
      o.line("%s", adj_mul(in, j, d, e).c_str());
    }
    o.indent--;
    o.line("}");
    o.indent--;
    o.line("}");
  }

  void win_adj_a(int i) {
    const TInstr& in = P.ins[i];
    const int a = clamp_off(in.a, 8 * nt), b = clamp_off(in.b, 8 * nt);
    const int d = in.dst;
    o.line("// adjoint of window %d, phase A: per-output coefficients, chunk scans -> shared", i);
    o.line("using BW%d = WinOp<SumOp<%d>, %d, %d, NT>;", i, kw(), -b, -a);
    o.line("static_assert(BW%d::FLOATS == %d, \"host / device scratch layout mismatch\");", i, win_floats(kw(), -b, -a, nt));
    o.line("Vec<%d> k%d[8], ksf%d[8], kpf%d[8];", kw(), i, i, i);
    if (P.sem == SEM_SMOOTH) o.line("unsigned lost%d = 0;", i);
    o.line("BW%d::fill_guards(smem + %d, c.tid);", i, bscr_off[i]);
    o.line("if (live) {");
    o.indent++;
    if (P.sem == SEM_SMOOTH) {
      o.line("lost%d = smooth_adj_coeff<%s, %d, %d>(c, g%d, acc%d, k%d);", i, mode_name(in.mode), a, b, d, i, i);
      o.line("if (lost%d) reinterpret_cast<volatile int*>(smem + %d)[%d] = 1;", i, flag_off, i);
    } else {
      o.line("gm_adj_coeff<SEM, %s, %d, %d>(c, rp, g%d, v%d, k%d);", mode_name(in.mode), a, b, d, d, i);
    }
    o.line("BW%d::phase1(smem + %d, c.tid, k%d, ksf%d, kpf%d);", i, bscr_off[i], i, i, i);
    o.indent--;
    o.line("} else {");
    o.indent++;
    o.line("BW%d::phase1_dead(smem + %d, c.tid);", i, bscr_off[i]);
    o.indent--;
    o.line("}");
  }

  void win_adj_b(int i) {
    const TInstr& in = P.ins[i];
    const int a = clamp_off(in.a, 8 * nt), b = clamp_off(in.b, 8 * nt);
    const bool neg = in.negmask & 1;
    const int s = in.src[0], d = in.dst;
    o.line("// adjoint of window %d, phase B: window sum of the coefficients x per-input factor", i);
    o.line("if (live) {");
    o.indent++;
    o.line("Vec<%d> w[8];", kw());
    o.line("BW%d::phase2(smem + %d, c.tid, k%d, ksf%d, kpf%d, w);", i, bscr_off[i], i, i, i);
    if (P.sem == SEM_SMOOTH) {
      if (relift()) {  // experiment: the lifted operand is not kept in registers, it is lifted again from the stash
        o.line("V ts[8];");
        o.line("{ float tt[8]; load8(smem + %d, c.tid, tt); _Pragma(\"unroll\") for (int i = 0; i < 8; ++i) ts[i] = V{{tt[i]}}; }", stash_off[i]);
        o.line("R%d::A xh[8];", i);
        o.line("lift8<R%d, VW>(c, rp%d, ts, xh);", i, i);
        o.line("smooth_adj_apply<%s, %s>(w, xh, g%d);", neg ? "true" : "false", may_nan[s] ? "true" : "false", s);
      } else {
        o.line("smooth_adj_apply<%s, %s>(w, x%d, g%d);", neg ? "true" : "false", may_nan[s] ? "true" : "false", i, s);
      }
    }
    else o.line("gm_adj_apply<SEM, %s, %s, %s>(c, rp, w, v%d, g%d);", mode_name(in.mode), neg ? "true" : "false", may_nan[s] ? "true" : "false", s, s);
    o.indent--;
    o.line("}");
    if (P.sem == SEM_SMOOTH) {
      o.line("if (reinterpret_cast<volatile int*>(smem + %d)[%d]) {  // rare: exact scatter of the outputs whose shifted sum underflowed", flag_off, i);
      o.indent++;
      o.line("float* fix = smem + %d;", fix_off[i]);
      o.line("{ const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}; store8(fix, c.tid, z); }");
      o.line("simt::sync();");
      o.line("if (lost%d) smooth_adj_scatter<%s, %d, %d, NT>(c, rp, lost%d, g%d, v%d, smem + %d, fix);", i, mode_name(in.mode), a, b, i, d, d,
             stash_off[i]);
      o.line("simt::sync();");
      o.line("{ float t[8]; load8(fix, c.tid, t); _Pragma(\"unroll\") for (int i = 0; i < 8; ++i) g%d[i] %s t[i]; }", s, neg ? "-=" : "+=");
      o.line("simt::sync();");
      o.indent--;
      o.line("}");
    }
  }

  // ---------------------------------------------------------------- cumulative backward
  std::string adj_add(int src, bool neg, const char* expr) const {  // g_src += expr, through a Not
    char b[256];
    if (neg) snprintf(b, sizeof(b), "{ const V t_ = %s; g%d[i].x[0] -= t_.x[1]; g%d[i].x[1] -= t_.x[0]; }", expr, src, src);
    else snprintf(b, sizeof(b), "{ const V t_ = %s; g%d[i].x[0] += t_.x[0]; g%d[i].x[1] += t_.x[1]; }", expr, src, src);
    return b;
  }
  void cum_adjoints() {
    const int L = 8 * nt;
    for (int st = n_stages - 1; st >= 0; --st) {
      o.line("// ================= adjoint stage %d", st);
      for (int i = P.n_instr - 1; i >= 0; --i) {
        const TInstr& in = P.ins[i];
        if (ins_level[i] != st || is_staged(in)) continue;
        const int n = in.nsrc, d = in.dst;
        o.line("// adjoint of the n-ary %s %d: the cotangent goes to the first extremal operand", in.mode == MODE_AND ? "min" : "max", i);
        o.line("if (live) {");
        o.indent++;
        o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i) {");
        o.indent++;
        std::string a = "const V a[" + std::to_string(n) + "] = {";
        for (int j = 0; j < n; ++j) a += (j ? ", " : "") + operand(in, j, "i");
        o.line("%s};", a.c_str());
        o.line("V ga[%d];", n);
        o.line("cum_nary_bwd<%s, %d>(a, g%d[i], ga);", mode_name(in.mode), n, d);
        for (int j = 0; j < n; ++j) {
          char e[32];
          snprintf(e, sizeof(e), "ga[%d]", j);
          o.line("%s", adj_add(in.src[j], (in.negmask >> j) & 1, e).c_str());
        }
        o.indent--;
        o.line("}");
        o.indent--;
        o.line("}");
      }
      for (int i = P.n_instr - 1; i >= 0; --i) {
        const TInstr& in = P.ins[i];
        if (ins_level[i] != st || !is_staged(in)) continue;
        const int d = in.dst;
        if (cum_lockstep(in)) {
          const bool nl = in.negmask & 1, nr = in.negmask & 2;
          const int OWN = cum_sp[i], HL = cum_sp[i] + L, HR = cum_sp[i] + 2 * L, CM = cum_sp[i] + 3 * L;
          o.line("// adjoint of until %d: lockstep shared-tail walk per channel (hand-off planes, no atomics)", i);
          o.line("float la%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
          o.line("float lb%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
          o.line("float ra%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
          o.line("float rb%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
          o.line("simt::sync();  // (the planes below are the forward scratch of this level: its last readers may still be in phase 2)");
          for (int w = 0; w < 2; ++w)
            o.line("cum_until_adj_stage<NT>(c, U%d::lplane(smem + %d, %d), smem + %d);", i, scr_off[i], w, CM + w * roundup(nt, 32));
          o.line("simt::sync();");
          for (int w = 0; w < 2; ++w) {
            o.line("{");
            o.line("  float gw[8];");
            o.line("  _Pragma(\"unroll\") for (int i = 0; i < 8; ++i) gw[i] = g%d[i].x[%d];", d, w);
            o.line("  cum_until_adj_walk<%d, %d, NT>(c, live, U%d::lplane(smem + %d, %d), U%d::rplane(smem + %d, %d), smem + %d, smem + %d, smem + %d, smem + %d, gw, %s%d, %s%d);",
                   in.a, in.b, i, scr_off[i], w, i, scr_off[i], w, CM + w * roundup(nt, 32), HL, HR, OWN, w ? "lb" : "la", i, w ? "rb" : "ra", i);
            o.line("}");
          }
          o.line("if (live) {");
          o.indent++;
          o.line("cum_add_ab8<%s>(la%d, lb%d, g%d);", nl ? "true" : "false", i, i, (int)in.src[0]);
          o.line("cum_add_ab8<%s>(ra%d, rb%d, g%d);", nr ? "true" : "false", i, i, (int)in.src[1]);
          o.indent--;
          o.line("}");
        } else if (is_until(in)) {
          const bool nl = in.negmask & 1, nr = in.negmask & 2;
          o.line("// adjoint of until %d: the r-routes are gathered into registers, the l-routes scattered with merged runs", i);
          o.line("plane_zero8(smem + %d, c.tid);", sp_off);
          o.line("plane_zero8(smem + %d, c.tid);", sp_off + L);
          o.line("stash8<2>(smem + %d, c.tid, g%d, 0);  // the cotangent of the until, for the gather (zeros past the horizon)", sp_off + 2 * L, d);
          o.line("stash8<2>(smem + %d, c.tid, g%d, 1);", sp_off + 3 * L, d);
          o.line("simt::sync();");
          o.line("float ra%d[8], rb%d[8];", i, i);
          o.line("if (live) {");
          o.indent++;
          for (int w = 0; w < 2; ++w)
            o.line("cum_until_gather_r<%d, %d, NT>(c, g%d, %d, U%d::lplane(smem + %d, %d), U%d::rplane(smem + %d, %d), smem + %d, %s%d);", in.a, in.b, d,
                   w, i, scr_off[i], w, i, scr_off[i], w, sp_off + (2 + w) * L, w ? "rb" : "ra", i);
          for (int w = 0; w < 2; ++w)
            o.line("cum_until_scatter_l<%d, %d, NT>(c, g%d, %d, U%d::lplane(smem + %d, %d), U%d::rplane(smem + %d, %d), smem + %d);", in.a, in.b, d, w,
                   i, scr_off[i], w, i, scr_off[i], w, sp_off + w * L);
          o.indent--;
          o.line("}");
          o.line("simt::sync();");
          o.line("if (live) {");
          o.indent++;
          o.line("cum_accumulate8<%s>(smem + %d, smem + %d, c.tid, g%d);", nl ? "true" : "false", sp_off, sp_off + L, (int)in.src[0]);
          o.line("cum_add_ab8<%s>(ra%d, rb%d, g%d);", nr ? "true" : "false", i, i, (int)in.src[1]);
          o.indent--;
          o.line("}");
          o.line("simt::sync();");
        } else if (arg_window(in)) {
          const bool neg = in.negmask & 1;
          o.line("// adjoint of window %d: scatter to the recorded arg%s", i, in.mode == MODE_AND ? "min" : "max");
          const int SP = cum_sp[i];  // (the window's own forward planes: its results are in registers by now)
          o.line("simt::sync();  // (the planes below were read by the previous adjoint / the forward sweep of this level)");
          o.line("plane_zero8(smem + %d, c.tid);", SP);
          o.line("plane_zero8(smem + %d, c.tid);", SP + L);
          o.line("simt::sync();");
          o.line("if (live) {");
          o.indent++;
          o.line("arg_scatter8(smem + %d, acca%d, g%d, 0);", SP, i, d);
          o.line("arg_scatter8(smem + %d, accb%d, g%d, 1);", SP + L, i, d);
          o.indent--;
          o.line("}");
          o.line("simt::sync();");
          o.line("if (live) cum_accumulate8<%s>(smem + %d, smem + %d, c.tid, g%d);", neg ? "true" : "false", SP, SP + L, (int)in.src[0]);
          o.line("simt::sync();");
        } else {  // Eventually: a sum
          const int a = clamp_off(in.a, 8 * nt), b = clamp_off(in.b, 8 * nt);
          const bool neg = in.negmask & 1;
          o.line("// adjoint of the sum window %d: backward-looking window sum of the cotangent", i);
          o.line("using BW%d = WinOp<SumOp<2>, %d, %d, NT>;", i, -b, -a);
          o.line("static_assert(BW%d::FLOATS == %d, \"host / device scratch layout mismatch\");", i, win_floats(2, -b, -a, nt));
          o.line("Vec<2> ksf%d[8], kpf%d[8];", i, i);
          o.line("BW%d::fill_guards(smem + %d, c.tid);", i, bscr_off[i]);
          o.line("if (live) BW%d::phase1(smem + %d, c.tid, g%d, ksf%d, kpf%d);", i, bscr_off[i], d, i, i);
          o.line("else BW%d::phase1_dead(smem + %d, c.tid);", i, bscr_off[i]);
          o.line("simt::sync();");
          o.line("if (live) {");
          o.indent++;
          o.line("Vec<2> w[8];");
          o.line("BW%d::phase2(smem + %d, c.tid, g%d, ksf%d, kpf%d, w);", i, bscr_off[i], d, i, i);
          o.line("cum_add8<%s>(w, g%d);", neg ? "true" : "false", (int)in.src[0]);
          o.indent--;
          o.line("}");
          o.line("simt::sync();");
        }
      }
    }
  }

  void wwin_adj(int i) {
    const TInstr& in = P.ins[i];
    const bool neg = in.negmask & 1;
    o.line("// adjoint of the weighted window %d: correlation of the output coefficients with the weights", i);
    o.line("using WA%d = WeightedWinAdj<SEM, %s, %d, %d, NT>;", i, mode_name(in.mode), in.a, in.b);
    o.line("if (live) WA%d::phase1(smem + %d, c, rp, P.wbuf + %d, g%d, v%d);", i, bscr_off[i], (int)in.w_off, (int)in.dst, (int)in.dst);
    o.line("simt::sync();");
    o.line("if (live) WA%d::phase2<%s>(smem + %d, c, rp, P.wbuf + %d, v%d, g%d);", i, neg ? "true" : "false", bscr_off[i], (int)in.w_off,
           (int)in.src[0], (int)in.src[0]);
    o.line("simt::sync();");
  }

  void until_adj_smooth(int i) {
    const TInstr& in = P.ins[i];
    const int L = 8 * nt;
    const bool nl = in.negmask & 1, nr = in.negmask & 2;
    if (P.sem == SEM_SMOOTH) {
      const int GL = hand_off[i], GR = hand_off[i] + L, LA = aux_off[i], RH = aux_off[i] + L, ag_off = aux_off[i] + 2 * L, d = in.dst;
      o.line("// adjoint of until %d: shared-tail walk downwards in the shifted linear domain, chunks in lockstep (no atomics)", i);
      o.line("float uc%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
      o.line("float ugl%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
      o.line("float ugr%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
      o.line("unsigned ulost%d = 0;", i);
      o.line("smooth_until_adj_stage<NT>(c, rp%d, U%d::lplane(smem + %d, 0), U%d::rplane(smem + %d, 0), smem + %d, smem + %d, smem + %d);",
             i, i, scr_off[i], i, scr_off[i], LA, RH, ag_off);
      o.line("if (live) {");
      o.indent++;
      o.line("ulost%d = smooth_until_adj_coeff<%d, %d>(c, g%d, ua%d, uc%d);", i, in.a, in.b, d, i, i);
      o.line("if (ulost%d) reinterpret_cast<volatile int*>(smem + %d)[%d] = 1;", i, flag_off, i);
      o.indent--;
      o.line("}");
      o.line("simt::sync();");
      o.line("smooth_until_adj_walk<%d, %d, NT>(c, live, smem + %d, smem + %d, smem + %d, smem + %d, smem + %d, uc%d, ugl%d, ugr%d);",
             in.a, in.b, LA, RH, ag_off, GL, GR, i, i, i);
      o.line("if (live) smooth_until_adj_finish<%s, %s>(smem + %d, smem + %d, c.tid, ugl%d, ugr%d, g%d, g%d);", nl ? "true" : "false",
             nr ? "true" : "false", LA, RH, i, i, (int)in.src[0], (int)in.src[1]);
      o.line("if (reinterpret_cast<volatile int*>(smem + %d)[%d]) {  // rare: exact log-domain scatter of the outputs whose A left the healthy range",
             flag_off, i);
      o.indent++;
      o.line("simt::sync();");
      o.line("plane_zero8(smem + %d, c.tid);", GL);
      o.line("plane_zero8(smem + %d, c.tid);", GR);
      o.line("simt::sync();");
      o.line("if (ulost%d) smooth_until_scatter<%d, %d, NT>(c, rp, ulost%d, g%d, U%d::lplane(smem + %d, 0), U%d::rplane(smem + %d, 0), smem + %d, smem + %d);",
             i, in.a, in.b, i, d, i, scr_off[i], i, scr_off[i], GL, GR);
      o.line("simt::sync();");
      o.line("if (c.p0 < c.nlive) {");
      o.line("  accumulate8<%s>(smem + %d, c.tid, g%d);", nl ? "true" : "false", GL, (int)in.src[0]);
      o.line("  accumulate8<%s>(smem + %d, c.tid, g%d);", nr ? "true" : "false", GR, (int)in.src[1]);
      o.line("}");
      o.indent--;
      o.line("}");
      o.line("simt::sync();");
      return;
    }
    if (P.sem == SEM_AGM && lockstep_adj(in)) {  // agm, uniform weights
      const int G0 = hand_off[i], G1 = hand_off[i] + L, GR = hand_off[i] + 2 * L, AN = aux_off[i], AL = aux_off[i] + roundup(nt, 32), d = in.dst;
      o.line("// adjoint of until %d: shared-tail walk downwards, chunks in lockstep (no atomics)", i);
      o.line("float uk%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
      o.line("float uz0%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
      o.line("float uz1%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
      o.line("float urr%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
      o.line("unsigned upos%d = 0;", i);
      o.line("agm_until_adj_stage<NT>(c, U%d::lplane(smem + %d, 0), smem + %d, smem + %d);", i, scr_off[i], AN, AL);
      o.line("if (live) upos%d = agm_until_adj_coeff<%d, %d>(c, g%d, v%d, uk%d);", i, in.a, in.b, d, d, i);
      o.line("simt::sync();");
      o.line("agm_until_adj_walk<%d, %d, NT>(c, live, P.ins[%d].w[0], P.ins[%d].w[1], U%d::lplane(smem + %d, 0), U%d::rplane(smem + %d, 0), smem + %d, smem + %d,",
             in.a, in.b, i, i, i, scr_off[i], i, scr_off[i], AN, AL);
      o.line("    smem + %d, smem + %d, smem + %d, uk%d, upos%d, uz0%d, uz1%d, urr%d);", G0, G1, GR, i, i, i, i, i);
      o.line("if (live) agm_until_adj_finish<%s, %s>(U%d::lplane(smem + %d, 0), U%d::rplane(smem + %d, 0), c.tid, P.ins[%d].w[0], P.ins[%d].w[1], uz0%d, uz1%d, urr%d, g%d, g%d);",
             nl ? "true" : "false", nr ? "true" : "false", i, scr_off[i], i, scr_off[i], i, i, i, i, i, (int)in.src[0], (int)in.src[1]);
      o.line("simt::sync();  // the hand-off planes go back to the window scratch");
      return;
    }
    int GLo = sp_off, GRo = sp_off + L;  // scatter planes of the per-output sweep
    const bool dg_fast = P.sem == SEM_DGMSR && lockstep_adj(in);
    if (dg_fast) {
      const int G0 = hand_off[i], G1 = hand_off[i] + L, GR = hand_off[i] + 2 * L, A0 = aux_off[i], A1 = aux_off[i] + roundup(nt, 32), d = in.dst;
      GLo = G0;
      GRo = G1;
      o.line("if (P.ins[%d].w[0] == 1.f && P.ins[%d].w[1] == 1.f) {  // unit weights_pair: lockstep shared-tail walk, closed forms per case", i, i);
      o.indent++;
      o.line("float ukA%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
      o.line("float ukG%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
      o.line("float uz0%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
      o.line("float uz1%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
      o.line("float urr%d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};", i);
      o.line("dgmsr_until_adj_stage<NT>(c, U%d::lplane(smem + %d, 0), smem + %d, smem + %d);", i, scr_off[i], A0, A1);
      o.line("if (live) dgmsr_until_adj_coeff<%d, %d>(c, rp, g%d, v%d, ukA%d, ukG%d);", in.a, in.b, d, d, i, i);
      o.line("simt::sync();");
      o.line("dgmsr_until_adj_walk<%d, %d, NT>(c, live, rp, U%d::lplane(smem + %d, 0), U%d::rplane(smem + %d, 0), smem + %d, smem + %d,", in.a, in.b, i,
             scr_off[i], i, scr_off[i], A0, A1);
      o.line("    smem + %d, smem + %d, smem + %d, ukA%d, ukG%d, uz0%d, uz1%d, urr%d);", G0, G1, GR, i, i, i, i, i);
      o.line("if (live) dgmsr_until_adj_finish<%s, %s>(U%d::lplane(smem + %d, 0), U%d::rplane(smem + %d, 0), c.tid, uz0%d, uz1%d, urr%d, g%d, g%d);",
             nl ? "true" : "false", nr ? "true" : "false", i, scr_off[i], i, scr_off[i], i, i, i, (int)in.src[0], (int)in.src[1]);
      o.line("simt::sync();  // the hand-off planes go back to the window scratch");
      o.indent--;
      o.line("} else {");
      o.indent++;
    }
    o.line("// adjoint of until %d: per-output sweeps, scatter through shared memory", i);
    o.line("plane_zero8(smem + %d, c.tid);", GLo);
    o.line("plane_zero8(smem + %d, c.tid);", GRo);
    o.line("simt::sync();");
    {
      char wl[48] = "nullptr", wr[48] = "nullptr";
      if (in.w_off >= 0) snprintf(wl, sizeof(wl), "P.wbuf + %d", (int)in.w_off);
      if (in.w2_off >= 0) snprintf(wr, sizeof(wr), "P.wbuf + %d", (int)in.w2_off);
      o.line("if (live) gm_until_scatter<SEM, %d, %d, NT>(c, rp, P.ins[%d].w[0], P.ins[%d].w[1], %s, %s, g%d, U%d::lplane(smem + %d, 0),", in.a,
             in.b, i, i, wl, wr, (int)in.dst, i, scr_off[i]);
      o.line("    U%d::rplane(smem + %d, 0), smem + %d, smem + %d);", i, scr_off[i], GLo, GRo);
    }
    o.line("simt::sync();");
    o.line("if (live) {");
    o.indent++;
    o.line("accumulate8<%s>(smem + %d, c.tid, g%d);", nl ? "true" : "false", GLo, (int)in.src[0]);
    o.line("accumulate8<%s>(smem + %d, c.tid, g%d);", nr ? "true" : "false", GRo, (int)in.src[1]);
    o.indent--;
    o.line("}");
    o.line("simt::sync();");
    if (dg_fast) {
      o.indent--;
      o.line("}");
    }
  }

  std::string backward() {
    plan();
    plan_bounds();
    plan_nan();
    emit_header("backward (value and gradient)");
    o.line("PF_KERNEL(NT, %d) %s(const PF_GRID_CONSTANT TileProgram P, const PF_GRID_CONSTANT TileGeom g,", minb, kernel_name().c_str());
    o.line("    const float* __restrict__ signal, const float* __restrict__ grad_out, float scale, float* __restrict__ grad_signal,");
    o.line("    float* __restrict__ out, double* __restrict__ mean_acc) {");
    o.indent = 1;
    std::vector<char> has(kMaxSlots, 0);
    has[P.root] = 1;
    for (int i = 0; i < P.n_instr; ++i) {
      has[P.ins[i].dst] = 1;
      for (int j = 0; j < P.ins[i].nsrc; ++j) has[P.ins[i].src[j]] = 1;
    }
    for (int q = 0; q < P.n_preds; ++q) has[P.preds[q].dst] = 1;
    if (!src_bwd() && !cum_bwd()) {
      for (int i = 0; i < P.n_instr; ++i)
        if (P.ins[i].op == T_NARY && P.ins[i].dst == P.root) fold_root = i;
      for (int sl = 0; sl < kMaxSlots; ++sl)
        // (-0.f: x + (-0) == x for every x, so the first accumulation into an adjoint folds to a move)
        if (has[sl]) o.line("float g%d[8] = {-0.f, -0.f, -0.f, -0.f, -0.f, -0.f, -0.f, -0.f};", sl);
    }
    emit_forward_body();
    o.line("// ================= value (optional) and cotangent");
    o.line("float part = 0.f;  // (its block sum joins the fused mean at the very end: a barrier there costs nothing)");
    if (src_bwd()) o.line("if (live) part = out ? store_root_src<%s>(c, g.T, v%d, out) : 0.f;", P.root_neg ? "true" : "false", (int)P.root);
    else o.line("if (live) part = out ? store_root<SEM, %s>(c, g.T, v%d, out) : 0.f;", P.root_neg ? "true" : "false", (int)P.root);
    if (src_bwd()) {
      // Classic: every output routes its cotangent to ONE leaf sample (the arg-min / arg-max chain of
      // torch.min / max(dim), first index on ties): scatter into one shared plane per predicate
      o.line("// ================= one-hot routing of the root cotangent (classic.py:25-68 under autograd)");
      if (!src_image()) {
        for (int q = 0; q < P.n_preds; ++q) o.line("plane_zero8(smem + %d, c.tid);", sp_off + q * 8 * nt);
        o.line("simt::sync();");
        o.line("if (live) classic_scatter<%s, NT>(c, g.T, grad_out, scale, v%d, smem + %d);", P.root_neg ? "true" : "false", (int)P.root, sp_off);
        o.line("simt::sync();");
        for (int q = 0; q < P.n_preds; ++q) {
          o.line("float g%d[8];", (int)P.preds[q].dst);
          o.line("load8(smem + %d, c.tid, g%d);", sp_off + q * 8 * nt, (int)P.preds[q].dst);
        }
      } else {
      o.line("zero_grad_image<NT, D>(smem + %d, c.tid);", sp_off);
      o.line("simt::sync();");
      o.line("if (live) classic_scatter_rows<%s, NT, D>(c, P, g.T, grad_out, scale, v%d, smem + %d);", P.root_neg ? "true" : "false", (int)P.root,
             sp_off);
      o.line("simt::fence_async_smem();");
      o.line("simt::sync();");
      o.line("store_grad_image<NT, D>(c, g.T, smem + %d, grad_signal);", sp_off);
      o.line("if (out && mean_acc) block_sum_to<NT>(c, part, mean_acc, smem + %d);", red_floats - 32);
      o.indent = 0;
      o.line("}");
      return o.s;
      }
    } else if (cum_bwd()) {
      for (int sl = 0; sl < kMaxSlots; ++sl)
        if (has[sl]) {
          o.line("V g%d[8];", sl);
          o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i) g%d[i] = V{{0.f, 0.f}};", sl);
        }
      o.line("if (live) root_cotangent2<%s>(c, g.T, grad_out, scale, g%d);", P.root_neg ? "true" : "false", (int)P.root);
      cum_adjoints();
    } else {
      if (fold_root < 0) o.line("if (live) root_cotangent<%s>(c, g.T, grad_out, scale, g%d);", P.root_neg ? "true" : "false", (int)P.root);
      for (int st = n_stages - 1; st >= 0; --st) {
        o.line("// ================= adjoint stage %d", st);
        for (int i = P.n_instr - 1; i >= 0; --i)
          if (ins_level[i] == st && !is_staged(P.ins[i])) nary_bwd(i);
        for (int i = P.n_instr - 1; i >= 0; --i)
          if (ins_level[i] == st && is_until(P.ins[i])) until_adj_smooth(i);
        for (int i = P.n_instr - 1; i >= 0; --i)
          if (ins_level[i] == st && is_wwin(P.ins[i])) wwin_adj(i);
        bool any = false;
        for (int i = P.n_instr - 1; i >= 0; --i)
          if (ins_level[i] == st && is_window(P.ins[i])) {
            win_adj_a(i);
            any = true;
          }
        if (any) o.line("simt::sync();");
        for (int i = P.n_instr - 1; i >= 0; --i)
          if (ins_level[i] == st && is_window(P.ins[i])) win_adj_b(i);
      }
    }
    o.line("// ================= gradient of the signal: chain rule through the affine predicates");
    const bool tma_st = tma_store_capable();
    if (tma_st) o.line("float gr[8][D];");
    o.line("if (live) {");
    o.indent++;
    if (D <= 8) {
      if (!tma_st) o.line("float gr[8][D];");
      o.line("_Pragma(\"unroll\") for (int i = 0; i < 8; ++i) {");
      o.indent++;
      for (int d = 0; d < D; ++d) {
        std::string e;
        for (int q = 0; q < P.n_preds; ++q)
          for (int k = 0; k < P.preds[q].n_terms; ++k)
            if (P.preds[q].ch[k] == d) {
              char b[96];
              if (P.preds[q].pair)
                snprintf(b, sizeof(b), "%sg%d[i].x[%d]", e.empty() ? "" : " + ", (int)P.preds[q].dst, k);
              else if (cum_bwd())
                snprintf(b, sizeof(b), "%sP.preds[%d].coeff[%d] * cum_pred_adj(v%d[i], g%d[i])", e.empty() ? "" : " + ", q, k, (int)P.preds[q].dst,
                         (int)P.preds[q].dst);
              else
                snprintf(b, sizeof(b), "%sP.preds[%d].coeff[%d] * g%d[i]", e.empty() ? "" : " + ", q, k, (int)P.preds[q].dst);
              e += b;
            }
        o.line("gr[i][%d] = %s;", d, e.empty() ? "0.f" : e.c_str());
      }
      o.indent--;
      o.line("}");
      if (!tma_st) o.line("store_grad_rows<D>(c, g.T, gr, grad_signal);");
    } else {
      for (int d = 0; d < D; ++d) {
        std::string e;
        for (int q = 0; q < P.n_preds; ++q)
          for (int k = 0; k < P.preds[q].n_terms; ++k)
            if (P.preds[q].ch[k] == d) {
              char b[96];
              if (P.preds[q].pair)
                snprintf(b, sizeof(b), "%sg%d[i].x[%d]", e.empty() ? "" : " + ", (int)P.preds[q].dst, k);
              else if (cum_bwd())
                snprintf(b, sizeof(b), "%sP.preds[%d].coeff[%d] * cum_pred_adj(v%d[i], g%d[i])", e.empty() ? "" : " + ", q, k, (int)P.preds[q].dst,
                         (int)P.preds[q].dst);
              else
                snprintf(b, sizeof(b), "%sP.preds[%d].coeff[%d] * g%d[i]", e.empty() ? "" : " + ", q, k, (int)P.preds[q].dst);
              e += b;
            }
        o.line("{ float gx[8]; _Pragma(\"unroll\") for (int i = 0; i < 8; ++i) gx[i] = %s; store_grad_channel(c, g.T, D, %d, gx, grad_signal); }",
               e.empty() ? "0.f" : e.c_str(), d);
      }
    }
    o.indent--;
    o.line("}");
    if (tma_st)
      o.line("store_grad_rows_tma<D, NT>(c, g, live, gr, grad_signal, smem + %d, part, out ? mean_acc : nullptr, smem + %d);", stage_off,
             red_floats - 32);
    else
      o.line("if (out && mean_acc) block_sum_to<NT>(c, part, mean_acc, smem + %d);", red_floats - 32);
    o.indent = 0;
    o.line("}");
    return o.s;
  }
};

// ---------------------------------------------------------------------- post-pass: one `if (live)` per barrier interval
// The emitters above write one `if (live) { ... } [else { ... }]` per operator.  Between two
// barriers those blocks are independent straight-line code; separate blocks cost a
// BSSY / BRA / BSYNC triple each and, worse, end the basic block, so ptxas cannot interleave the
// dependent add chains of two windows.  This pass rewrites every barrier interval as
//     <declarations and all-thread statements, in order>
//     if (live) { {body 1} {body 2} ... } else { dead 1  dead 2 ... }
// Only lines that are provably order-independent of the live bodies are hoisted (declarations,
// uniform scalars derived from the block-wide bounds, guard-band fills); anything else is a fence.
bool starts_with(const std::string& s, const char* p) { return s.compare(0, strlen(p), p) == 0; }
bool ends_with(const std::string& s, const char* p) {
  const size_t n = strlen(p);
  return s.size() >= n && s.compare(s.size() - n, n, p) == 0;
}
std::string strip_comment(const std::string& l) {
  const size_t c = l.find("//");
  std::string t = c == std::string::npos ? l : l.substr(0, c);
  while (!t.empty() && t.back() == ' ') t.pop_back();
  return t;
}
bool hoistable(const std::string& raw) {
  const std::string l = strip_comment(raw);
  if (l.size() <= 2) return true;  // blank / comment-only line
  if (!starts_with(l, "  ") || l[2] == ' ' || !ends_with(l, ";")) return false;
  const std::string t = l.substr(2);
  if (starts_with(t, "using ") || starts_with(t, "static_assert(")) return true;
  if (starts_with(t, "RP rp") || (starts_with(t, "rp") && t.find(".shift = ") != std::string::npos)) return true;
  if (t.find("::fill_guards(") != std::string::npos && t.find("if ") == std::string::npos) return true;
  if (starts_with(t, "unsigned lost") && ends_with(t, " = 0;")) return true;
  if (t == "float part = 0.f;") return true;
  // array declarations: no initialiser, or a brace list of zeros
  if (t.find('(') == std::string::npos && t.find('[') != std::string::npos) {
    const size_t eq = t.find('=');
    if (eq == std::string::npos) return true;
    const std::string init = t.substr(eq + 1);
    return init.find_first_not_of(" {}-0.f,;") == std::string::npos;
  }
  return false;
}
std::string merge_live_blocks(const std::string& src) {
  std::vector<std::string> lines;
  for (size_t i = 0; i < src.size();) {
    const size_t e = src.find('\n', i);
    if (e == std::string::npos) {
      lines.push_back(src.substr(i));
      break;
    }
    lines.push_back(src.substr(i, e - i));
    i = e + 1;
  }
  std::string out;
  std::vector<std::vector<std::string>> live, dead;
  auto flush = [&]() {
    if (live.empty()) return;
    const bool many = live.size() > 1;
    out += "  if (live) {\n";
    for (auto& b : live) {
      if (many) out += "   {\n";
      for (auto& l : b) out += l + "\n";
      if (many) out += "   }\n";
    }
    bool any_dead = false;
    for (auto& b : dead) any_dead |= !b.empty();
    if (any_dead) {
      out += "  } else {\n";
      for (auto& b : dead)
        for (auto& l : b) out += l + "\n";
    }
    out += "  }\n";
    live.clear();
    dead.clear();
  };
  bool in_body = false, seen_kernel = false;
  for (size_t i = 0; i < lines.size(); ++i) {
    const std::string& l = lines[i];
    if (!in_body) {
      out += l + "\n";
      if (starts_with(l, "PF_KERNEL(")) seen_kernel = true;
      if (seen_kernel && ends_with(l, "{")) in_body = true;
      continue;
    }
    if (l == "}") {  // end of the kernel
      flush();
      out += l + "\n";
      in_body = seen_kernel = false;
      continue;
    }
    const std::string t = strip_comment(l);
    if (t == "  if (live) {") {
      std::vector<std::string> b, d;
      size_t j = i + 1;
      bool in_dead = false, ok = false;
      for (; j < lines.size(); ++j) {
        const std::string u = strip_comment(lines[j]);
        if (u == "  } else {") {
          in_dead = true;
          continue;
        }
        if (u == "  }") {
          ok = true;
          break;
        }
        (in_dead ? d : b).push_back(lines[j]);
      }
      if (!ok) {  // not understood: leave the rest alone
        flush();
        for (; i < lines.size(); ++i) out += lines[i] + "\n";
        return out;
      }
      live.push_back(b);
      dead.push_back(d);
      i = j;
      continue;
    }
    if (starts_with(t, "  if (live) ") && ends_with(t, ";") && t.find('{') == std::string::npos) {
      live.push_back({"    " + t.substr(strlen("  if (live) "))});
      dead.push_back({});
      continue;
    }
    if (hoistable(l)) {
      out += l + "\n";
      continue;
    }
    // fence: a top-level statement or block this pass does not reorder around
    flush();
    out += l + "\n";
    if (starts_with(l, "  ") && l.size() > 2 && l[2] != ' ' && ends_with(t, "{")) {
      for (++i; i < lines.size(); ++i) {
        out += lines[i] + "\n";
        if (strip_comment(lines[i]) == "  }") break;
      }
    }
  }
  return out;
}

}  // namespace

// Tile geometry.  A CTA of nt threads stages 8*nt samples; forward tiles carry the formula's
// look-ahead H as a right halo, backward tiles carry it on both sides.  Every multiple of 32
// threads is a candidate: the cost is the number of samples staged per trace (halo + the idle
// tail of the last tile), with a mild preference for CTAs of at most 256 threads (more resident
// CTAs per SM hide each other's barriers).
int fused_geometry(const Lowered& lw, int64_t B, int64_t T, int32_t D, int force_nt, bool backward, TileGeom* g, char* msg,
                   int msg_len) {
  auto fail = [&](int code, const char* what) {
    if (msg && msg_len > 0) snprintf(msg, msg_len, "%s", what);
    return code;
  };
  if (B < 0 || T < 0 || D < 1) return fail(PYSTL_E_INVALID, "bad shape");
  if (lw.lookahead < 0) return fail(PYSTL_E_UNSUPPORTED, "open-ended look-ahead");
  if (T >= (int64_t(1) << 31) - 65536) return fail(PYSTL_E_UNSUPPORTED, "horizon beyond 2^31 samples (tile arithmetic is 32-bit)");
  const TileProgram& P = lw.prog;
  for (int q = 0; q < P.n_preds; ++q)
    for (int k = 0; k < P.preds[q].n_terms; ++k)
      if (P.preds[q].ch[k] < 0 || P.preds[q].ch[k] >= D) return fail(PYSTL_E_INVALID, "predicate reads a channel the signal does not have");
  const int H = lw.lookahead;
  // (two-channel signals: 16, so that the first row a tile stores starts a 128-byte line of the TMA staging area)
  const int hh = backward ? roundup(H, D == 2 ? 16 : 8) : 0;
  const int64_t Tn = std::max<int64_t>(T, 1);
  auto build = [&](int nt, TileGeom* out) -> bool {
    TileGeom t{};
    t.B = B;
    t.T = T;
    t.D = D;
    t.nt = nt;
    t.lext = 8 * nt;
    t.vw = P.sem == SEM_CUMULATIVE ? 2 : 1;
    t.bwd = backward ? 1 : 0;
    t.hh = hh;
    t.n_slots = P.n_slots;
    if (Tn <= t.lext) {
      t.lout = t.lext;
      t.n_tiles = 1;
    } else {
      t.lout = backward ? t.lext - 2 * hh : ((t.lext - H) / 8) * 8;
      if (t.lout < 8 || t.lout * 4 < t.lext) return false;
      t.n_tiles = (int)((Tn + t.lout - 1) / t.lout);
    }
    *out = t;
    return true;
  };
  if (force_nt > 0) {
    if (force_nt % 32 != 0 || force_nt > kMaxThreads) return fail(PYSTL_E_UNSUPPORTED, "forced CTA size is not a multiple of 32");
    if (!build(force_nt, g)) return fail(PYSTL_E_UNSUPPORTED, "forced CTA size does not fit the formula's look-ahead");
    return PYSTL_OK;
  }
  double best = 0;
  bool have = false;
  for (int nt = 64; nt <= kMaxThreads; nt += 32) {
    if (nt > 512 && nt % 128 != 0) continue;
    TileGeom t{};
    if (!build(nt, &t)) continue;
    double cost = (double)t.n_tiles * t.lext;
    if (nt > 256) cost *= 1.0 + 0.12 * (nt - 256) / 256.0;
    if (nt < 128) cost *= 1.05;
    // (measured, cfg2 value + gradient: 3 tiles of 192 threads at 80 registers, 4 CTAs per SM: 1.602 ms; 2 tiles of 288 threads at
    // <= 75 registers with the operands re-lifted, 3 CTAs per SM = 27 warps: 1.555 ms -- same padded work)
    if (backward && nt == kWideSmoothThreads && t.n_tiles > 1 && wide_smooth_backward(P)) cost = (double)t.n_tiles * t.lext * 0.99;
    if (!have || cost < best * 0.995) {
      have = true;
      best = cost;
      *g = t;
    }
  }
  if (!have) return fail(PYSTL_E_UNSUPPORTED, "look-ahead does not fit a tile");
  return PYSTL_OK;
}

int fused_generate(const Lowered& lw, int D, int nt, int minb, bool backward, FusedSource* out, char* msg, int msg_len) {
  Gen g(lw, D, nt, minb, backward);
  std::string why;
  if (!g.supported(&why)) {
    if (msg && msg_len > 0) snprintf(msg, msg_len, "%s", why.c_str());
    return PYSTL_E_UNSUPPORTED;
  }
  out->source = backward ? g.backward() : g.forward();
  // (measured on B200: the merged form is 1-6 % faster except for the agm / dgmsr value-and-gradient
  // kernels, where the longer basic blocks cost registers: cfg3 3.19 -> 3.27 ms)
  const bool merge = !backward || lw.prog.sem == SEM_SMOOTH || lw.prog.sem == SEM_CUMULATIVE;
  const char* me = getenv("PYSTL_B200_MERGE");  // 0 / 1: override (experiments)
  if (me ? atoi(me) != 0 : merge) out->source = merge_live_blocks(out->source);
  out->name = g.kernel_name();
  out->smem_floats = g.smem_floats;
  out->nt = nt;
  out->minb = minb;
  return PYSTL_OK;
}

}  // namespace pystl
