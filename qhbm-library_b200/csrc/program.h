// program.h -- the compiled "sweep / pass / op" program shared by the host
// compiler (plan.cpp), the CUDA kernels (kernels.cu) and the test-only schedule
// verifier (tests/native/verify_plan.cpp).  Plain-old-data only.
//
// Vocabulary
//   state    complex64[2^n] amplitudes of one unique bitstring's circuit
//   tile     2^T amplitudes of one state that one CTA holds in shared memory
//   sweep    one pass of every tile of every state through shared memory
//            (global -> smem -> [passes] -> global); the tile <-> state bit map is
//            chosen per sweep so that the gates it runs act inside the tile
//   pass     smem -> registers -> [ops] -> smem with K "register qubits": each
//            thread holds the 2^K amplitudes that differ only in those K bits
//   op       one fused gate block (or gradient inner product) applied in registers
#pragma once
#include <stdint.h>

namespace qhbm {

constexpr int kMaxRegQubits = 5;
constexpr int kMaxTileQubits = 14;
constexpr int kMaxQubits = 30;
constexpr int kConstGroupBits = 7;  // width of one "thread-constant" phase table
constexpr int kMaxRuns = 16;
constexpr int kStageOps = 192;     // ops of one pass: the compiler closes a pass before its program exceeds the
constexpr int kStageCoef = 1792;   // shared-memory staging buffer (op descriptors / coefficient floats)
constexpr int kStageOpsAdj = 256;  // the same for a gradient pass of the adjoint kernel (ops + reduction tasks)
constexpr int kGaccFloats = 1024;  // reduced gradient sums (one flush window) kept in shared memory

enum OpType : int32_t {
  OP_NOP = 0,
  OP_MAT1,        // p0: register position; coef -> 2x2 complex (8 floats)
  OP_MAT2,        // p0: 0 -> positions (1,0), 1 -> positions (3,2); coef -> 4x4 (32 floats),
                  //     matrix index = 2*bit(hi position) + bit(lo position)
  OP_DCONST_TAB,  // F *= tab[(gidx >> aux0) & aux1]; coef -> complex table
  OP_DCONST_PAIR, // F *= c[2*bit(aux0) + bit(aux1)] (aux1 < 0: c[bit(aux0)]); coef -> 4 complex
  OP_DREG_TAB,    // amp[r] *= F * tab[r]; F = 1; coef -> 2^K entries of 4 floats (re, im, -im, im);
                  //     aux0 = 0: F is known to be 1
  OP_DAPPLY,      // amp[r] *= F; F = 1
  OP_DCROSS,      // amps with register bit p0 = v: *= c[2*bit(aux0) + v]; coef -> 4 complex
  OP_XROT,        // (c I - i s X) on register position p0, global phase dropped; coef -> (c, s)
  OP_YROT,        // (c I - i s Y) on register position p0, global phase dropped; coef -> (c, s)
  // gradient ops (adjoint sweeps only); value lands in scratch slot gslot
  OP_GRAD_MAT1,   // 2 Re <lam| M |psi>, M 2x2 at coef, position p0
  OP_GRAD_MAT2,   // same for 4x4, p0 as in OP_MAT2
  OP_XROTM,       // several OP_XROT in one op: p0 = mask of register positions, coef -> K x (c, s, kappa, -),
                  //     aux0 = mask of positions with a gradient; slot of position P: byte P of aux1
                  //     (P < 4) or p1 (P = 4)
  OP_YROTM,       // same for OP_YROT
  OP_XROTF,       // OP_XROTM applied to ALL K positions without tests (inactive ones hold the identity
                  //     rotation): no branches, no register reconciliation at merge points.  When the 4th
                  //     coefficient of position 0 is non-zero, positions 0..K-2 hold tan(theta) and are applied
                  //     as 1-FMA-per-component unnormalised rotations; position K-1 restores the scale
  OP_GRAD_X,      // kappa * Im <lam| X_p0 |psi>; coef -> kappa   (two-level X-type gate)
  OP_GRAD_Y,      // kappa * Im <lam| Y_p0 |psi>; coef -> kappa
  OP_GD_BEGIN,    // starts a run of aux0 diagonal-gradient ops (they share conj(lam)*psi), sorted by kind:
                  //     p0 OP_GD_CONST, then p1 OP_GD_REG1 / OP_GD_MIX, then the OP_GD_REG2;
                  //     aux1 != 0: pair marginals are needed (some OP_GD_REG2 in the run).
                  //     Device form (lower_device_program): coef = mask of the marginal vectors the run needs
                  //     (bit 0: T, bit 1 + p: S[p], bit 1 + K + pi: SS[pi]); gslot = first scratch unit.  The
                  //     thread stores those vectors and skips the aux0 descriptors, which only the host reads
  OP_GD_CONST,    // entries M[sel] at coef (complex); sel = 2*bit(aux0)+bit(aux1) (aux1<0: bit(aux0))
  OP_GD_REG1,     // sel = register bit p0; coef -> 2 complex
  OP_GD_REG2,     // sel = 2*regbit(p0) + regbit(p1); coef -> 4 complex
  OP_GD_MIX,      // one register bit p0 and one constant bit aux0; coef -> 4 complex indexed
                  //     [2*bit(aux0) + regbit(p0)]
  // observable passes (DevPass::ngrad == -1): h = H psi accumulated in the lambda registers
  OP_HX,          // h[r] += s * tab[r] * psi[r ^ p0]: p0 = register xor mask with one or two bits set,
                  //     coef -> 2^K real coefficients, s = (-1)^{parity(gbase & aux0)} (aux0 = 0: +1)
  OP_HD,          // h[r] += s * tab[r] * psi[r]  (diagonal terms), same fields
  // reduction tasks of a gradient pass (device program only; they follow the pass's executed ops)
  OP_TASK_F,      // gacc[coef] = sum over the CTA's threads of the float vector in scratch unit p0
  OP_TASK_C,      // gacc[coef], gacc[coef + 1] = sum over the threads with (tid & aux0) == aux0 of the complex
                  //     vector in scratch units p0, p0 + 1
};

struct DevOp {  // 32 bytes
  int32_t type;
  int32_t p0, p1;
  int32_t coef;   // float offset into the coefficient buffer
  int32_t gslot;  // gradient scratch slot within the pass, or -1
  int32_t aux0, aux1;
  int32_t pad;
};

// Device form of a DevOp: 16 bytes, one shared-memory load per op.  type, p0, p1 and gslot are bytes
// (-1 becomes 255; no kernel path tests them for sign).
struct PackedOp {
  uint32_t w0;  // type | p0 << 8 | p1 << 16 | gslot << 24
  int32_t coef, aux0, aux1;
};
inline PackedOp pack_op(const DevOp& o) {
  PackedOp q;
  q.w0 = ((uint32_t)o.type & 0xffu) | (((uint32_t)o.p0 & 0xffu) << 8) | (((uint32_t)o.p1 & 0xffu) << 16) |
         (((uint32_t)o.gslot & 0xffu) << 24);
  q.coef = o.coef;
  q.aux0 = o.aux0;
  q.aux1 = o.aux1;
  return q;
}

// One gradient value of a flush window: how to combine reduced sums into d<H>/d(symbol).
//   kind 0: the float sum gacc[i_tot] (rotation / matrix-block gradients are complete per thread)
//   kind 1: one-qubit diagonal gate, M = diag(m0, m1) at coef: U1 = A, U0 = tot - A
//   kind 2: two-qubit diagonal gate, M entries indexed 2 * bitA + bitB: U11 = AB, U10 = A - AB,
//           U01 = B - AB, U00 = tot - A - B + AB
// tot / A / B / AB are complex sums (u, -v) of w = conj(lam) psi = u + i v at gacc[i_*], taken over the
// tile (tot), over the amplitudes with bit A set, with bit B set, with both set.  A bit outside the tile
// is a condition on the CTA's tile offset instead (cond_*: state-index bit, -1: none): the marginal is
// the sum named by the index if the bit is set in goff and zero otherwise.
// value = 2 sum_sel (Re m_sel * U_sel.x + Im m_sel * U_sel.y)
struct DevGradDesc {  // 32 bytes
  int32_t kind;
  int32_t sym;
  int32_t coef;
  int16_t i_tot, i_a, i_b, i_ab;
  int8_t cond_a, cond_b;
  int8_t pad0[2];
  int32_t pad1[2];
};

struct DevPass {  // 224 bytes
  int32_t regbit[kMaxRegQubits];     // tile-local bit of register position j
  int32_t sorted[kMaxRegQubits];     // the same bits in ascending order
  int32_t op_begin, op_end;
  int32_t ngrad, gsym_off;           // gradient slots of this pass -> symbols gsym[gsym_off..];
                                     // ngrad == -1 marks an observable pass (OP_HX / OP_HD only)
  int32_t coef_begin, coef_end;      // float range of the coefficient buffer this pass reads
  uint32_t eoff8[1 << kMaxRegQubits]; // swizzled smem BYTE offset of register amplitude r (XOR with the thread base)
  int32_t next_op_end, next_coef_end;  // ends of the NEXT pass's ranges (they start where this pass's end):
                                       // what the kernel needs to prefetch that program while this pass runs
  // device program only (HostPlan::dev_passes):
  int32_t exec_end;                    // ops [op_begin, exec_end) run per thread, [exec_end, op_end) are tasks
  int32_t gd_flush_begin, gd_flush_end;  // gradient descriptors evaluated (and added to the float64
  int32_t pad[3];                        // accumulators) after this pass: the end of a flush window
};

// Contiguous run of tile-local bits mapped to contiguous state-index bits.
struct BitRun {
  int8_t local_start, global_start, len, pad;
};

enum LaunchFlags : uint32_t {
  LF_INIT_BASIS = 1u << 0,  // tile := |basis> restricted to the tile (no load)
  LF_LOAD_PSI = 1u << 1,
  LF_LOAD_LAM = 1u << 2,
  LF_EXPECT = 1u << 3,      // expectation values (+ lambda = sum_j g_j H_j psi if adjoint)
  LF_STORE_PSI = 1u << 4,
  LF_STORE_LAM = 1u << 5,
  LF_WRITE_STATE = 1u << 6, // debug: store psi into a caller buffer
  LF_PSI_ALT = 1u << 7,     // psi is stored into the alternate workspace buffer (other CTAs of this launch
                            // still read the old psi across tiles); later launches load from there
  // After the first forward sweep the state is zero outside the tile that held the basis state.  That sweep
  // does not store its all-zero tiles (their CTAs exit at once) and the next sweep does not load them: an
  // amplitude whose out-of-first-tile index bits (`sparse_mask`) differ from the basis index is zero-filled
  // without a memory access.  Halves the DRAM traffic of a forward computation with two sweeps.
  LF_SPARSE_OUT = 1u << 8,
  LF_SPARSE_IN = 1u << 9,
};

// One kernel launch = one sweep of one chunk of states.
struct LaunchDesc {
  uint32_t flags;
  int32_t pass_a_begin, pass_a_end;  // passes run before the expectation phase (psi only)
  int32_t pass_b_begin, pass_b_end;  // passes run after it (psi and lambda, with gradients)
  int32_t n_runs, n_oruns;
  BitRun runs[kMaxRuns];             // tile-local bits -> state bits
  BitRun oruns[kMaxRuns];            // tile-id bits -> state bits (the out-of-tile bits)
  uint32_t tile_mask;                // state-index bits covered by the tile
  int32_t pass_h_begin, pass_h_end;  // observable passes run at the start of the expectation phase
  int32_t expect_stage;              // LF_EXPECT: which stage's tables this launch uses
  int32_t gslot_begin, gslot_count;  // gradient slots (indices into gsym) of all passes of this launch (host view)
  int32_t rng_begin, rng_end;        // its slice of the observable ranges (DevOpRange)
  int32_t grp_begin, grp_end;        // that stage's slice of the group / term tables (staged in shared memory)
  int32_t term_begin, term_end;
  uint32_t sparse_mask;              // LF_SPARSE_IN: state-index bits that must equal the basis index's
  // the thread's m-th tile element is local index (m << (T-K)) | tid:
  uint32_t moff[1 << kMaxRegQubits]; // its state-index contribution scatter(m << (T-K))
  uint16_t soff[1 << kMaxRegQubits]; // its swizzled smem contribution swz(m << (T-K))
};

// Pauli-sum tables for the expectation phase.
struct DevTerm {  // coefficient(i) += (kr + i ki) * (-1)^{parity(i & z)}
  float kr, ki;
  uint32_t z;
  uint32_t mword;      // bit m = parity(z & index bits of the thread's m-th amplitude) in the
                       // expectation launch (contiguous tile, amplitude m = m * nthreads + tid)
};
struct DevTermGroup {  // terms sharing one x-mask: H psi[i] += coefficient(i) * psi[i ^ x]
  uint32_t x;          // state-index xor mask
  int32_t xl;          // tile-local xor mask if the partner is inside the tile, else -1
  int32_t term_begin, term_end;  // terms with z != 0
  float k0r, k0i;      // sum of the z == 0 terms (index-independent part of the coefficient)
  int32_t is_complex;  // some term has an imaginary coefficient (odd number of Y)
  int32_t pad;
};
struct DevOpRange {  // the x-groups of one observable in one expectation stage; only observables
  int32_t group_begin, group_end;  // that have generic groups (or observable passes to finish) get one
  int32_t op, pad;
};
// Diagonal (Z-string) term evaluated through a Walsh-Hadamard transform of |psi|^2 (many-shard case)
struct DevDiagTerm {
  float coeff;
  uint32_t z;
  int32_t op;
  int32_t pad;
};
constexpr int kWhtMinTerms = 32;  // below this the direct evaluation is cheaper

// Coefficient-preparation jobs (run on the device once per call, from the symbols).
enum PrepKind : int32_t {
  PJ_MAT1 = 0,   // product of list gates (application order); a: dagger
  PJ_MAT2,       // single gate; a: dagger; b: swap qubit roles
  PJ_GRAD1,      // M = dG G^dagger for (gate list[0], param c)
  PJ_GRAD2,      // same, 4x4; b: swap qubit roles
  PJ_GDIAG,      // diagonal of M for a diagonal gate; b: swap roles (4 entries) ; 1q gate: 2 entries
  PJ_DTAB,       // phase table with 2^d entries; list = triples (gate, posA, posB): entry[v] =
                 //   prod_g diag_g[2*bit(v,posA) + bit(v,posB)] (posB < 0 -> 1q gate); a: dagger;
                 //   b: entries are 4 floats (re, im, -im, im) instead of (re, im)
  PJ_DPAIR,      // 4 (or 2) diagonal entries of one gate; a: dagger; b: swap roles
  PJ_ROT,        // (cos, sin)(pi t / 2) of an XPow / YPow gate; a: dagger (sin negated)
  PJ_KAPPA,      // kappa of a two-level X/Y-type gate (b: 0 = X, 1 = Y) for param c:
                 //   dG G^dagger = i(c0 I - kappa/2 A)  =>  d<H>/ds = kappa Im<lam|A|psi>
  PJ_PHASE,      // product of the dropped global phases e^{i pi t (g + 1/2)} of the list gates
  PJ_ROTF,       // all K rotations of one OP_XROTF: list = gate per register position (-1: identity);
                 //   a: dagger; b: mask of positions with a gradient.  Positions 0..K-2 are stored in the
                 //   unnormalised form (tan, -, kappa', 1) and the last one absorbs the product of their
                 //   cosines, unless some cosine is too small (then the plain (c, s, kappa, 0) form)
  PJ_NONE,       // removed job
  PJ_CONST,      // symbol-independent floats: the list holds their bit patterns
};
struct PrepJob {  // 32 bytes
  int32_t kind;
  int32_t out;        // float offset into the coefficient buffer
  int32_t a, b, c, d;
  int32_t list_off, list_len;
};

}  // namespace qhbm
