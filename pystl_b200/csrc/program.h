// Tile program: what a formula is lowered to before it reaches the kernel.
//
// A CTA evaluates the whole formula over one tile of one trace.  Every node
// value lives in a shared-memory "plane" (one float per sample of the tile,
// two for cumulative's (pos, neg)); a thread owns 8 consecutive samples of
// every plane.  The program is a flat list of instructions over plane slots,
// passed to the kernel by value (__grid_constant__), so decoding is a uniform
// constant-bank read.
#pragma once
#if defined(__CUDACC_RTC__)
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
#else
#include <stdint.h>
#endif

namespace pystl {

constexpr int kChunk = 8;       // samples per thread (two float4)
constexpr int kMaxSrc = 16;     // operands of one n-ary instruction
constexpr int kMaxTerms = 8;    // == PYSTL_MAX_TERMS
constexpr int kMaxPreds = 32;   // distinct affine predicates per formula
constexpr int kMaxInstr = 64;   // lowered instructions per formula (the kernel parameter block stays below 16 KB of the 32 KB limit)
constexpr int kMaxThreads = 1024;
constexpr int kMaxSlots = kMaxPreds + kMaxInstr + 8;  // value slots of one formula
constexpr int kMaxWeights = 1024;  // window-position weights of one formula (agm / dgmsr, fused tier)

enum Sem : int32_t {
  SEM_CLASSIC = 0,
  SEM_SMOOTH = 1,
  SEM_CUMULATIVE = 2,
  SEM_AGM = 3,
  SEM_DGMSR = 4,
  // kernel-internal: classic values that carry the leaf they came from (predicate, position,
  // sign) through every min / max; the classic backward pass is one scatter through them
  SEM_CLASSIC_SRC = 5
};

// which reduction an instruction performs
enum Mode : int32_t {
  MODE_AND = 0,  // conjunction / ALWAYS
  MODE_OR = 1,   // disjunction / EVENTUALLY
  MODE_SUM = 2   // cumulative EVENTUALLY (cumulative.py:61-72)
};

enum TOp : int32_t {
  T_NARY = 0,        // dst[t] = reduce_i src_i[t]                       (And / Or)
  T_WIN_DIRECT = 1,  // dst[t] = reduce_{k in [a,b]} src[t+k], b-a+1 <= 8
  T_WIN_VH = 2,      // same, van Herk / Gil-Werman over 8-sample chunks, b-a+1 >= 9
  T_WIN_SUFFIX = 3,  // dst[t] = reduce_{k >= a} src[t+k] up to the horizon (tile must reach T)
  T_UNTIL = 4,       // dst[t] = OR_{i in [a,b]} AND(AND_{k<=i} l[t+k], r[t+i])   O(b) per sample
  T_CLAMP_SCAN = 5,  // classic unbounded until: W(t) = min(l_t, max(r_t, W(t+1)))
  T_WIN_WEIGHTED = 6 // agm / dgmsr window with per-position weights (agm.py:18-35): O(W) per sample, fused tier only
};

struct TPred {
  int16_t dst;  // plane slot
  int16_t n_terms;
  int16_t f64;  // classic only: evaluate in float64 and round once (bit-exact against the float64 reference)
  int16_t pair; // cumulative only: the value is (signal[ch[0]], signal[ch[1]]) as it stands (a host-computed cumulative value)
  int16_t ch[kMaxTerms];
  float coeff[kMaxTerms];
  float bias;
};

struct TInstr {
  int16_t op;    // TOp
  int16_t mode;  // Mode
  int16_t dst;   // plane slot of the result
  int16_t nsrc;
  int16_t src[kMaxSrc];  // plane slots
  uint16_t negmask;      // bit i: operand i is negated (Not folded into the consumer)
  int16_t unbounded;     // T_UNTIL: b is the horizon
  uint16_t adj_first;    // backward: bit i: this instruction is the first to write operand i's adjoint
  int32_t a, b;          // window offsets relative to t (signed: the backward pass looks back)
  float w[kMaxSrc];      // n-ary weights; T_UNTIL: w[0], w[1] = weights_pair
  float sw;              // sum of the n-ary weights
  // offsets into TileProgram::wbuf, -1 = uniform: window weights / until weights_left, weights_right
  int16_t w_off, w2_off;
};

// Scalars derived from the semantics' kwargs, computed once on the host (lower.cpp).
struct RP {
  float k;         // smooth: log2(e) / tau
  float invk;      // smooth: tau * ln 2
  float shift;     // smooth: block-wide max of the exponent, set per instruction
  float eps;       // dgmsr
  float sqrt_eps;  // dgmsr
  float ln_eps;    // dgmsr
  float eps_p;     // dgmsr eps^p
  float inv2p;     // dgmsr 1 / (2p)
  int32_t p;
};

struct TileProgram {
  int32_t sem;      // Sem
  int32_t p;        // dgmsr exponent
  float tau;        // smooth temperature
  float eps;        // dgmsr eps
  int32_t n_preds;
  int32_t n_instr;
  int32_t n_slots;  // plane slots in use
  int16_t root;     // plane slot holding the formula value
  int16_t root_neg; // formula is Not(...) at the top
  int32_t start_slack;  // outputs at t > T-1-start_slack are NaN
  int32_t raises_always;
  RP rp;
  // backward only: plane holding the adjoint of value slot s (-1: none), and how many planes
  // (values + adjoints, after reuse of dead ones) the tile needs
  int16_t adj_of[kMaxSlots];
  int32_t n_planes_bwd;
  TPred preds[kMaxPreds];
  TInstr ins[kMaxInstr];
  int32_t n_wbuf;
  float wbuf[kMaxWeights];  // window-relative weights (truncated to the clipped window, agm.py:28-35)
};

// Geometry of one launch.
struct TileGeom {
  int64_t B, T;
  int32_t D;
  int32_t nt;        // threads per CTA
  int32_t lext;      // samples staged per tile = 8 * nt
  int32_t lout;      // outputs per tile (multiple of 8)
  int32_t n_tiles;   // tiles per trace
  int32_t vw;        // words per value (2 for cumulative)
  int32_t scr_planes;  // scratch planes
  int32_t bwd;         // geometry of the backward kernel: halo of `hh` samples on both sides
  int32_t hh;          // look-ahead rounded up to 8 (backward only)
  int32_t n_slots;     // value slots; the backward kernel adds one adjoint plane per slot
  int32_t gl, gr;      // scratch guard bands (floats, multiples of 8) left / right of the tile:
                       // hold the identity so that window reads need no bounds checks
  // shared-memory carve-up, in floats
  int32_t off_planes, off_scratch, off_chunk, off_red, smem_floats;
  // fused tier, D == 2: the signal viewed as rows of 64 bytes, fetched tile by tile with one TMA
  // tensor copy (SWIZZLE_64B) when the launch qualifies (pystl_b200.cu: fused_tma_setup)
  int32_t use_tma;
  // ... and grad_signal as [trace][64-byte row][16 floats] for the TMA store of a tile's gradient rows
  // (clipped at the end of the trace by the tensor bounds): st_nbox boxes of st_rows rows each
  int32_t use_tma_st, st_rows, st_nbox;
  alignas(64) unsigned long long tmap[16];     // CUtensorMap, opaque here
  alignas(64) unsigned long long tmap_st[16];
};

// Lowering result (host only).
struct Lowered {
  TileProgram prog;
  int32_t lookahead;  // -1: unbounded
  int32_t out_channels;
};

}  // namespace pystl

struct pystl_formula;
namespace pystl {
// backward: every node keeps its own plane (values are re-read by the adjoint pass) and
// constructs without a backward kernel are rejected
int lower_formula(const ::pystl_formula* f, bool backward, Lowered* out, char* msg, int msg_len);
// ssa: no plane-slot reuse (the fused tier keeps every value in its own registers)
int lower_formula_ex(const ::pystl_formula* f, bool backward, bool ssa, Lowered* out, char* msg, int msg_len);
// force_nt > 0 pins the CTA size (tests); 0 = choose
int plan_geometry(const Lowered& lw, int64_t B, int64_t T, int32_t D, int force_nt, bool backward, TileGeom* g, char* msg,
                  int msg_len);
}  // namespace pystl
