// sim_kernels.cuh -- sm_100a kernels for the per-bitstring state-vector path:
// basis state -> parameterised circuit -> PauliSum expectations -> adjoint gradient.
//
// Replaces the arithmetic of TFQ's TfqSimulateExpectation / TfqAdjointGradient ops
// (reached from /root/reference/qhbmlib/inference/qnn.py:134-138).  Design in DESIGN.md.
//
// One CTA owns one TILE (2^T amplitudes, complex64) of one state in shared memory.
// A PASS moves the tile smem -> registers with K "register qubits": each thread
// holds the 2^K amplitudes that differ only in those K index bits, applies every
// fused gate block scheduled for the pass in registers, and writes back.  The smem
// layout is nibble-XOR swizzled so that every pass is bank-conflict free for any
// contiguous choice of register bits.  No tensor cores: there is no dense contraction.
#pragma once
#include <cuda_pipeline.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gate_math.h"
#include "program.h"

namespace qhbm {

struct KernelArgs {
  LaunchDesc L;
  const DevPass* passes;
  const PackedOp* ops;
  const float* coef;
  const DevGradDesc* gdescs;  // gradient descriptors (flush windows of the gradient passes)
  const DevTerm* terms;
  const DevTermGroup* groups;
  const DevOpRange* opranges;
  const DevDiagTerm* dterms;  // diagonal terms evaluated through the WHT path
  int32_t n_dterms;
  int32_t n_groups, n_terms;  // sizes of the group / term tables (for staging them in shared memory)
  float2* psi;            // [chunk][2^n] workspace (multi-tile only)
  float2* lam;            // [chunk][2^n] workspace (multi-tile adjoint only)
  const uint64_t* basis;  // [chunk]
  const float* dgrad;     // [chunk, O] upstream gradients (adjoint) or nullptr
  double* eacc;           // [chunk, O] expectation accumulators
  double* gacc;           // [rows, P] gradient accumulators
  float2* psi_out;        // where LF_STORE_PSI writes (== psi unless LF_PSI_ALT)
  float2* state_out;      // debug
  int32_t n, T, O, P;
  int32_t grow0;          // accumulator row of this chunk's first state (per-state gradients)
  int32_t per_state;
  int32_t phase_coef;     // coef offset of the dropped global phase (debug state output), or -1
  int32_t async_tile;     // tiles are loaded with cp.async (default; QHBM_SYNC_TILE=1 turns it off)
  int32_t bulk_stage;     // pass programs are staged by the bulk-copy engine (cp.async.bulk + mbarrier; default;
                          // QHBM_NO_BULK_STAGE=1 restores the per-thread 16-byte copies)
  uint32_t coef_stride;   // floats between the coefficient tables of consecutive states: 0 = one table shared by
                          // all states (symbols f32[P]); else a multiple of 4 (symbols f32[U,P], one table per state)
};

// ---- bulk asynchronous copies (TMA engine, 1-D) completed through an mbarrier ----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

__device__ __forceinline__ uint32_t swz(uint32_t x) {
  return (x & ~15u) | ((x ^ (x >> 4) ^ (x >> 8) ^ (x >> 12)) & 15u);
}
__device__ __forceinline__ uint32_t scatter_bits(uint32_t l, const BitRun* runs, int nr) {
  uint32_t g = 0;
#pragma unroll 1
  for (int i = 0; i < nr; ++i)
    g |= ((l >> runs[i].local_start) & ((1u << runs[i].len) - 1u)) << runs[i].global_start;
  return g;
}
__device__ __forceinline__ uint32_t gather_bits(uint32_t g, const BitRun* runs, int nr) {
  uint32_t l = 0;
#pragma unroll 1
  for (int i = 0; i < nr; ++i)
    l |= ((g >> runs[i].global_start) & ((1u << runs[i].len) - 1u)) << runs[i].local_start;
  return l;
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// ---- packed FP32 (sm_100 FFMA2 / FMUL2 / FADD2: two lanes per instruction).  A complex amplitude is
// one 64-bit register pair (re, im); ptxas folds half swaps (.LO_HI), per-half sign patterns (.NP) and
// scalar broadcasts (.F32) into the operand modifiers of the packed instruction, so a complex
// multiply-add is 2 issue slots instead of 4.
__device__ __forceinline__ float2 bc(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 swp(float2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
// Sign pair in the constant bank.  A pair such as (-y, y) built with register moves is re-materialised by
// ptxas at every use (one MOV per packed instruction, measured); as the product bc(y) * kNP it is ONE
// FMUL2 whose result is an ordinary register pair.
static __constant__ float2 kNP = {-1.f, 1.f};  // (static: the header is compiled into two units)
__device__ __forceinline__ float2 np_pair(float y) { return mul2(bc(y), kNP); }  // (-y, y)
// A complex constant c prepared for packed use: re = c.x and q = (-c.y, c.y).
struct CK {
  float re;
  float2 q;
};
__device__ __forceinline__ CK make_ck(float2 c) { CK k; k.re = c.x; k.q = np_pair(c.y); return k; }
// a * c
__device__ __forceinline__ float2 cmulk(float2 a, const CK& k) { return fma2(k.q, swp(a), mul2(bc(k.re), a)); }
// acc + a * c
__device__ __forceinline__ float2 cfmak(float2 a, const CK& k, float2 acc) {
  return fma2(k.q, swp(a), fma2(bc(k.re), a, acc));
}
__device__ __forceinline__ float hsum(float2 a) { return a.x + a.y; }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Program data (ops, coefficients) is read through generic pointers: it lives in shared memory when the
// pass was staged there, in global memory otherwise.
__device__ __forceinline__ float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 ldg2(const float* p) { return *reinterpret_cast<const float2*>(p); }

// One op descriptor in registers: the 16-byte PackedOp as loaded.  type / p0 / p1 / gslot stay packed in
// w0 and are extracted where they are used (one instruction each), which keeps the interpreter's live
// state at 4 registers per descriptor instead of 7 -- the kernel runs at the 128-register cap.
struct OpRec {
  uint32_t w0;
  int coef, aux0, aux1;
  __device__ __forceinline__ int type() const { return (int)(w0 & 0xffu); }
  __device__ __forceinline__ int p0() const { return (int)((w0 >> 8) & 0xffu); }
  __device__ __forceinline__ int p1() const { return (int)((w0 >> 16) & 0xffu); }
  __device__ __forceinline__ int gslot() const { return (int)(w0 >> 24); }
};
__device__ __forceinline__ OpRec load_op(const PackedOp* op) {
  const int4 a = *reinterpret_cast<const int4*>(op);
  OpRec r;
  r.w0 = (uint32_t)a.x;
  r.coef = a.y; r.aux0 = a.z; r.aux1 = a.w;
  return r;
}

// ---------------------------------------------------------------------------------
// Register-level gate blocks.  P / LO are compile-time register positions.  BOTH
// applies the same block to psi (a) and lambda (b).
// ---------------------------------------------------------------------------------
template <int V>
struct IntC {
  static constexpr int value = V;
};
// Calls f(IntC<p>{}) with the register position as a compile-time constant.
template <int K, class F>
__device__ __forceinline__ void dispatch_pos(int p, F&& f) {
  switch (p) {
    case 0: f(IntC<0>{}); break;
    case 1: f(IntC<1>{}); break;
    case 2: f(IntC<2>{}); break;
    case 3: f(IntC<3>{}); break;
    default:
      if constexpr (K > 4) f(IntC<4>{});
      break;
  }
}

// Calls f(IntC<0>{}), ..., f(IntC<K-1>{}) in order (compile-time loop over register positions).
template <int K, int P = 0, class F>
__device__ __forceinline__ void for_each_pos(F&& f) {
  if constexpr (P < K) {
    f(IntC<P>{});
    for_each_pos<K, P + 1>(f);
  }
}

template <int K, int P>
__device__ __forceinline__ void mat1(float2 (&a)[1 << K], const float4 m0, const float4 m1) {
  const CK k00 = make_ck(make_float2(m0.x, m0.y)), k01 = make_ck(make_float2(m0.z, m0.w));
  const CK k10 = make_ck(make_float2(m1.x, m1.y)), k11 = make_ck(make_float2(m1.z, m1.w));
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) {
    if (r & (1 << P)) continue;
    const float2 x0 = a[r], x1 = a[r | (1 << P)];
    a[r] = cfmak(x1, k01, cmulk(x0, k00));
    a[r | (1 << P)] = cfmak(x1, k11, cmulk(x0, k10));
  }
}
// (c I - i s X): y0 = c x0 - i s x1, y1 = -i s x0 + c x1
template <int K, int P>
__device__ __forceinline__ void xrot(float2 (&a)[1 << K], const float c, const float s) {
  const float2 ms = np_pair(s);  // (-s, s)
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) {
    if (r & (1 << P)) continue;
    const float2 x0 = a[r], x1 = a[r | (1 << P)];
    a[r] = fma2(make_float2(-ms.x, -ms.y), swp(x1), mul2(bc(c), x0));
    a[r | (1 << P)] = fma2(make_float2(-ms.x, -ms.y), swp(x0), mul2(bc(c), x1));
  }
}
// unnormalised (I - i t X), t = tan: one FMA per component; the missing factor cos is restored later
template <int K, int P>
__device__ __forceinline__ void xrot_fast(float2 (&a)[1 << K], const float t) {
  const float2 mt = np_pair(t);  // (-t, t)
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) {
    if (r & (1 << P)) continue;
    const float2 x0 = a[r], x1 = a[r | (1 << P)];
    a[r] = fma2(make_float2(-mt.x, -mt.y), swp(x1), x0);
    a[r | (1 << P)] = fma2(make_float2(-mt.x, -mt.y), swp(x0), x1);
  }
}
// (c I - i s Y) = [[c, -s], [s, c]]
template <int K, int P>
__device__ __forceinline__ void yrot(float2 (&a)[1 << K], const float c, const float s) {
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) {
    if (r & (1 << P)) continue;
    const float2 x0 = a[r], x1 = a[r | (1 << P)];
    a[r] = fma2(bc(-s), x1, mul2(bc(c), x0));
    a[r | (1 << P)] = fma2(bc(s), x0, mul2(bc(c), x1));
  }
}
// Im <b| X_P |a> and Im <b| Y_P |a> over the thread's amplitudes
template <int K, int P>
__device__ __forceinline__ float im_bxa(const float2 (&a)[1 << K], const float2 (&b)[1 << K]) {
  // Im conj(b) a = b.x a.y - b.y a.x = the two halves of b * (a.y, -a.x), summed at the end
  float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) {
    if (r & (1 << P)) continue;
    const int q = r | (1 << P);
    s0 = fma2(b[r], make_float2(a[q].y, -a[q].x), s0);
    s1 = fma2(b[q], make_float2(a[r].y, -a[r].x), s1);
  }
  return hsum(add2(s0, s1));
}
template <int K, int P>
__device__ __forceinline__ float im_bya(const float2 (&a)[1 << K], const float2 (&b)[1 << K]) {
  float2 sp = make_float2(0.f, 0.f), sm = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) {
    if (r & (1 << P)) continue;
    const int q = r | (1 << P);
    sp = fma2(b[q], a[r], sp);
    sm = fma2(b[r], a[q], sm);
  }
  return hsum(sp) - hsum(sm);
}

// 2 Re sum_r conj(b_r) (M a)_r over the thread's amplitudes.
template <int K, int P>
__device__ __forceinline__ float grad_mat1(const float2 (&a)[1 << K], const float2 (&b)[1 << K],
                                           const float4 m0, const float4 m1) {
  const CK k00 = make_ck(make_float2(m0.x, m0.y)), k01 = make_ck(make_float2(m0.z, m0.w));
  const CK k10 = make_ck(make_float2(m1.x, m1.y)), k11 = make_ck(make_float2(m1.z, m1.w));
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) {
    if (r & (1 << P)) continue;
    const float2 x0 = a[r], x1 = a[r | (1 << P)];
    s = fma2(b[r], cfmak(x1, k01, cmulk(x0, k00)), s);
    s = fma2(b[r | (1 << P)], cfmak(x1, k11, cmulk(x0, k10)), s);
  }
  return 2.f * hsum(s);
}

// 4x4 block on register positions (LO+1, LO); matrix index = 2*bit(LO+1) + bit(LO).
// GRAD: returns 2 Re <b| M |a> and leaves a untouched.
template <int K, int LO, bool GRAD>
__device__ __forceinline__ float mat2(float2 (&a)[1 << K], const float2 (&b)[1 << K], const float* __restrict__ mp) {
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) {
    if (r & (3 << LO)) continue;
    float2 x[4], y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = a[r | (j << LO)];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // the matrix rows are re-read (uniform, staged in shared memory) instead of holding 32 registers
      const float4 m01 = ldg4(mp + 8 * i), m23 = ldg4(mp + 8 * i + 4);
      y[i] = cmulk(x[0], make_ck(make_float2(m01.x, m01.y)));
      y[i] = cfmak(x[1], make_ck(make_float2(m01.z, m01.w)), y[i]);
      y[i] = cfmak(x[2], make_ck(make_float2(m23.x, m23.y)), y[i]);
      y[i] = cfmak(x[3], make_ck(make_float2(m23.z, m23.w)), y[i]);
    }
    if constexpr (GRAD) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s = fma2(b[r | (i << LO)], y[i], s);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) a[r | (i << LO)] = y[i];
    }
  }
  return 2.f * hsum(s);
}

// amplitudes with register bit P = v are multiplied by e_v; exact identities are skipped
template <int K, int P>
__device__ __forceinline__ void mul_sel(float2 (&a)[1 << K], const float2 e0, const float2 e1, const bool id0,
                                        const bool id1) {
  const CK k0 = make_ck(e0), k1 = make_ck(e1);
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) {
    if (r & (1 << P)) { if (!id1) a[r] = cmulk(a[r], k1); }
    else { if (!id0) a[r] = cmulk(a[r], k0); }
  }
}

template <int N>
__device__ __forceinline__ float pick(const float (&v)[N], int i) {
  float r = v[0];
#pragma unroll
  for (int k = 1; k < N; ++k) r = (i == k) ? v[k] : r;
  return r;
}

// Marginal sums of w_r = conj(b_r) a_r = u_r + i v_r over the register index r, kept as (u, -v) pairs
// (so that Re(m w) = m.x u - m.y v is the half-sum of the plain product m * (u, -v)):
// totals T, per register bit S[p] (sum over r with bit p set) and per pair of register bits SS[pi]
// (both set; pi = ph (ph - 1) / 2 + pl, ph > pl).  No diagonal gate changes w, so one set of marginals
// serves a whole run of diagonal-gate gradients.  The sums are built four amplitudes at a time (a
// two-level subset-sum tree over register bits 0 and 1, then one accumulate per higher bit), every add a
// packed FADD2 on the (u, v) pair: ~45 adds instead of one per (amplitude, membership).
template <int K>
struct Marginals {
  static constexpr int NP = K * (K - 1) / 2;
  float2 T;
  float2 S[K];
  float2 SS[NP > 0 ? NP : 1];
};
__device__ __forceinline__ void acc2(float2& dst, const float2 v, const bool first) {
  dst = first ? v : add2(dst, v);
}
template <int K>
__device__ __forceinline__ void compute_marginals(const float2 (&a)[1 << K], const float2 (&b)[1 << K],
                                                  Marginals<K>& mg, const bool pairs) {
  static_assert(K >= 2, "register qubits");
  auto pidx = [](int ph, int pl) { return ph * (ph - 1) / 2 + pl; };
#pragma unroll
  for (int g = 0; g < (1 << (K - 2)); ++g) {
    float2 x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 4 * g + j;
      const float2 w = mul2(b[r], a[r]);                             // halves sum to  Re conj(b) a
      const float2 z = mul2(b[r], make_float2(-a[r].y, a[r].x));     // halves sum to -Im conj(b) a
      x[j] = make_float2(w.x + w.y, z.x + z.y);
    }
    const float2 s0 = add2(x[1], x[3]);   // register bit 0 set
    const float2 s1 = add2(x[2], x[3]);   // register bit 1 set
    const float2 t = add2(add2(x[0], x[1]), s1);
    acc2(mg.T, t, g == 0);
    acc2(mg.S[0], s0, g == 0);
    acc2(mg.S[1], s1, g == 0);
    if (pairs) acc2(mg.SS[pidx(1, 0)], x[3], g == 0);
#pragma unroll
    for (int p = 2; p < K; ++p) {
      if ((g >> (p - 2)) & 1) {
        const bool first = g == (1 << (p - 2));
        acc2(mg.S[p], t, first);
        if (pairs) {
          acc2(mg.SS[pidx(p, 0)], s0, first);
          acc2(mg.SS[pidx(p, 1)], s1, first);
        }
      }
    }
    if (pairs) {
#pragma unroll
      for (int ph = 3; ph < K; ++ph)
#pragma unroll
        for (int pl = 2; pl < ph; ++pl)
          if (((g >> (ph - 2)) & 1) && ((g >> (pl - 2)) & 1))
            acc2(mg.SS[pidx(ph, pl)], t, g == ((1 << (ph - 2)) | (1 << (pl - 2))));
    }
  }
}

template <int N>
__device__ __forceinline__ float2 pick2(const float2 (&v)[N], int i) {
  float2 r = v[0];
#pragma unroll
  for (int k = 1; k < N; ++k) {
    r.x = (i == k) ? v[k].x : r.x;
    r.y = (i == k) ? v[k].y : r.y;
  }
  return r;
}

// Diagonal-gate gradients of a run.  A diagonal gate with M = dG G^dagger = diag(m_sel) contributes
// 2 Re sum_i m_sel(i) w_i, i.e. it only needs the sums of w over the amplitudes that select each entry.
// The thread stores the marginal vectors the run's gates read (`vmask`: bit 0 T, bit 1 + p S[p], bit
// 1 + K + pi SS[pi]; complex, two scratch units each, in mask order from unit `unit0`); the pass's
// reduction tasks sum them over the CTA -- for gates on thread-index bits over the threads that have
// the bit set -- and the flush combines the sums per gate (DevGradDesc).  No per-gate work per thread.
template <int K>
__device__ __forceinline__ void store_marginals(const float2 (&a)[1 << K], const float2 (&b)[1 << K],
                                                const uint32_t vmask, const int unit0, float* scratch,
                                                const uint32_t tid, const uint32_t nthr) {
  Marginals<K> mg;
  compute_marginals<K>(a, b, mg, (vmask >> (1 + K)) != 0u);
  float2* v = reinterpret_cast<float2*>(scratch + (size_t)unit0 * nthr) + tid;
  *v = mg.T;
  v += nthr;
#pragma unroll
  for (int p = 0; p < K; ++p) {
    if ((vmask >> (1 + p)) & 1u) {
      *v = mg.S[p];
      v += nthr;
    }
  }
#pragma unroll
  for (int pi = 0; pi < Marginals<K>::NP; ++pi) {
    if ((vmask >> (1 + K + pi)) & 1u) {
      *v = mg.SS[pi];
      v += nthr;
    }
  }
}

// ---------------------------------------------------------------------------------
// One pass: smem tile -> registers -> ops -> smem tile.
// BOTH = false: forward (psi only).  BOTH = true: adjoint (psi and lambda, gradients).
// ---------------------------------------------------------------------------------
// Shared-memory state of a CTA besides the tiles.
//   stage   : two program buffers (pass descriptor + op descriptors + coefficients); while a pass runs from
//             one, the next pass's program streams into the other with cp.async (LDGSTS), so no pass waits
//             for global memory except the first of a range
//   gacc    : the launch's gradient sums of this CTA, flushed to the float64 accumulators once at the end
struct PassCtx {
  float4* stage;
  float* gacc;
  int buf;
  bool dbuf;    // two program buffers (adjoint kernel); the forward-only kernel keeps one, to fit three CTAs per SM
  int ops_cap;  // op descriptors per buffer: kStageOpsAdj in the adjoint kernel (ops + reduction tasks), else kStageOps
  uint64_t* bars;  // one mbarrier per program buffer (bulk staging)
  uint32_t phase;  // bit b: parity of the next completion of buffer b's mbarrier
  __device__ __forceinline__ int stride() const { return stage_f4(ops_cap); }
  __host__ __device__ static constexpr int stage_f4(int cap) { return (int)(sizeof(DevPass) / 16) + cap + kStageCoef / 4; }
};
constexpr int kPassF4 = (int)(sizeof(DevPass) / 16);
static_assert(sizeof(DevPass) % 16 == 0, "DevPass is copied in 16-byte pieces");

// Copies pass p's program into stage buffer `buf`.  ASYNC: cp.async, completed by stage_wait().
template <bool ASYNC>
__device__ __forceinline__ void stage_program(const KernelArgs& ka, float4* buf, const int ops_cap, const int p,
                                              const int op_begin, const int op_end, const uint32_t cb, const int n_cf) {
  const int tid = (int)threadIdx.x, nthr = (int)blockDim.x;
  const float4* g_ps = reinterpret_cast<const float4*>(ka.passes + p);
  const float4* g_ops = reinterpret_cast<const float4*>(ka.ops + op_begin);
  const float4* g_cf = reinterpret_cast<const float4*>(ka.coef + cb);  // coefficient slots are 16-byte aligned
  float4* s_ops = buf + kPassF4;
  float4* s_cf = buf + kPassF4 + ops_cap;
  const int n_ops = op_end - op_begin;
  if constexpr (ASYNC) {
    if (tid < kPassF4) __pipeline_memcpy_async(buf + tid, g_ps + tid, 16);
    for (int i = tid; i < n_ops; i += nthr) __pipeline_memcpy_async(s_ops + i, g_ops + i, 16);
    for (int i = tid; i < n_cf; i += nthr) __pipeline_memcpy_async(s_cf + i, g_cf + i, 16);
    __pipeline_commit();
  } else {
    if (tid < kPassF4) buf[tid] = __ldg(g_ps + tid);
    for (int i = tid; i < n_ops; i += nthr) s_ops[i] = __ldg(g_ops + i);
    for (int i = tid; i < n_cf; i += nthr) s_cf[i] = __ldg(g_cf + i);
  }
}

// The same copy issued by ONE thread to the bulk-copy engine: three cp.async.bulk (descriptor, ops,
// coefficients) that complete on the buffer's mbarrier.  Sizes are multiples of 16 bytes by construction.
__device__ __forceinline__ void stage_program_bulk(const KernelArgs& ka, float4* buf, uint64_t* bar, const int ops_cap,
                                                   const int p, const int op_begin, const int op_end,
                                                   const uint32_t cb, const uint32_t n_cf) {
  const uint32_t n_ops = (uint32_t)(op_end - op_begin);
  mbar_expect_tx(bar, (uint32_t)sizeof(DevPass) + 16u * (n_ops + n_cf));
  bulk_g2s(buf, ka.passes + p, (uint32_t)sizeof(DevPass), bar);
  if (n_ops) bulk_g2s(buf + kPassF4, ka.ops + op_begin, 16u * n_ops, bar);
  if (n_cf) bulk_g2s(buf + kPassF4 + ops_cap, ka.coef + cb, 16u * n_cf, bar);
}

// Start of a pass: makes pass p's program current (staging it now if it is the first of its range, else
// waiting for the prefetch), starts the prefetch of pass p + 1, and returns the views into the buffer.
struct PassView {
  const DevPass* ps;      // in shared memory
  const PackedOp* ops;    // indexed by absolute op number
  const float* coef;      // indexed by absolute float offset
  int op_begin, op_end;   // op_end = end of the per-thread ops (DevPass::exec_end); tasks follow up to task_end
  int task_end;
};
// cu: float offset of this state's coefficient table (0 unless the call has one row of symbols per state).
__device__ __forceinline__ PassView begin_pass(const KernelArgs& ka, PassCtx& cx, const int p, const bool first,
                                               const bool last, const uint32_t cu) {
  const bool bulk = ka.bulk_stage != 0;
  if (first || !cx.dbuf) {
    __syncthreads();  // the previous phase is done with the tiles and the stage buffers
    const DevPass* gp = ka.passes + p;
    if (bulk) {
      if (threadIdx.x == 0) {
        const int cb0 = __ldg(&gp->coef_begin), ce0 = __ldg(&gp->coef_end);
        stage_program_bulk(ka, cx.stage + cx.buf * cx.stride(), cx.bars + cx.buf, cx.ops_cap, p, __ldg(&gp->op_begin),
                           __ldg(&gp->op_end), cu + (uint32_t)cb0, (uint32_t)((ce0 - cb0 + 3) / 4));
      }
    } else {
      const int cb0 = __ldg(&gp->coef_begin), ce0 = __ldg(&gp->coef_end);
      stage_program<false>(ka, cx.stage + cx.buf * cx.stride(), cx.ops_cap, p, __ldg(&gp->op_begin), __ldg(&gp->op_end),
                           cu + (uint32_t)cb0, (ce0 - cb0 + 3) / 4);
    }
    if (ka.async_tile) __pipeline_wait_prior(0);  // cp.async tile loads of this thread (experiment switch)
  } else if (!bulk) {
    __pipeline_wait_prior(0);
  }
  if (bulk) {  // every thread observes the completion: the engine's writes are visible to it afterwards
    mbar_wait(cx.bars + cx.buf, (cx.phase >> cx.buf) & 1u);
    cx.phase ^= 1u << cx.buf;
  }
  __syncthreads();  // program visible; the previous pass's tile stores and gradient reduction are complete
  float4* buf = cx.stage + cx.buf * cx.stride();
  PassView v;
  v.ps = reinterpret_cast<const DevPass*>(buf);
  v.op_begin = v.ps->op_begin;
  v.op_end = v.ps->exec_end;
  v.task_end = v.ps->op_end;
  const int cb = v.ps->coef_begin;
  v.ops = reinterpret_cast<const PackedOp*>(buf + kPassF4) - v.op_begin;
  v.coef = reinterpret_cast<const float*>(buf + kPassF4 + cx.ops_cap) - cb;
  if (cx.dbuf) {
    if (!last) {  // passes of a range are consecutive: the next program starts where this one ends
      if (bulk) {
        if (threadIdx.x == 0)
          stage_program_bulk(ka, cx.stage + (cx.buf ^ 1) * cx.stride(), cx.bars + (cx.buf ^ 1), cx.ops_cap, p + 1,
                             v.task_end, v.ps->next_op_end, cu + (uint32_t)v.ps->coef_end,
                             (uint32_t)((v.ps->next_coef_end - v.ps->coef_end + 3) / 4));
      } else {
        stage_program<true>(ka, cx.stage + (cx.buf ^ 1) * cx.stride(), cx.ops_cap, p + 1, v.task_end, v.ps->next_op_end,
                            cu + (uint32_t)v.ps->coef_end, (v.ps->next_coef_end - v.ps->coef_end + 3) / 4);
      }
    }
    cx.buf ^= 1;
  }
  return v;
}

// GEN = false compiles the general-matrix ops out (OP_MAT1 / OP_MAT2 / OP_GRAD_MAT* / OP_YROTM): their 4x4
// blocks set the register allocation of the whole interpreter loop (120 bytes of spills per thread in the
// adjoint kernel, 16 without them), so plans that never use them -- X-power / diagonal-gate circuits such
// as the hardware-efficient ansatz -- run on the lean instantiation (HostPlan::lean).
template <int K, bool BOTH, bool GEN>
__device__ __forceinline__ void run_pass(const KernelArgs& ka, PassCtx& cx, const int p, const bool first,
                                         const bool last, float2* s_psi, float2* s_lam, uint32_t goff, uint32_t u) {
  constexpr int R = 1 << K;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  const PassView pv = begin_pass(ka, cx, p, first, last, u * ka.coef_stride);
  const DevPass* ps = pv.ps;
  const int op_begin = pv.op_begin, op_end = pv.op_end;
  const PackedOp* ops_base = pv.ops;
  const float* coef_base = pv.coef;
  uint32_t base = tid;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int sp = ps->sorted[j];
    base = ((base >> sp) << (sp + 1)) | (base & ((1u << sp) - 1u));
  }
  const uint32_t B = swz(base);
  const uint32_t gbase = goff | scatter_bits(base, ka.L.runs, ka.L.n_runs);
  float2 a[R];
  float2 b[BOTH ? R : 1];
  // Byte addressing: amplitude r sits at tile + (8 B ^ eoff8[r]) -- one LOP3 per access, the tile base folds
  // into the LDS / STS address operand.
  const uint32_t B8 = 8u * B;
  {
    const uint4* ep = reinterpret_cast<const uint4*>(ps->eoff8);
#pragma unroll
    for (int i = 0; i < R / 4; ++i) {
      const uint4 w = ep[i];
      const uint32_t eo[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        a[4 * i + r] = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(s_psi) + (B8 ^ eo[r]));
        if constexpr (BOTH)
          b[4 * i + r] = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(s_lam) + (B8 ^ eo[r]));
      }
    }
  }
  const int ngrad = BOTH ? ps->ngrad : 0;
  // Gradient scratch [slot][thread] = the psi and lambda tiles themselves, which are dead while every
  // amplitude sits in registers (dedicated scratch was measured: it costs 16 KB of shared memory per CTA,
  // which shrinks L1 to ~20 KB and turns the kernel's few register spills into L2 round trips).
  float* scratch = reinterpret_cast<float*>(s_psi);
  if (ngrad > 0) __syncthreads();  // every thread has read its amplitudes: the tiles become scratch

  float2 F = make_float2(1.f, 0.f);
  int oi = op_begin;
  while (oi < op_end) {
    const OpRec op = load_op(ops_base + oi);  // one LDS.128 from the staged program
    const float* cf = coef_base + op.coef;
    int step = 1;
    if (op.type() == OP_GD_BEGIN) step += op.aux0;
    switch (op.type()) {
      case OP_XROTM: {  // p0: positions that are rotated; aux0: positions with a gradient (a superset for the
        for_each_pos<K>([&](auto pc) {  // circuit's first gates, whose un-rotation nothing needs)
          constexpr int P = decltype(pc)::value;
          if ((op.p0() | (BOTH ? op.aux0 : 0)) & (1 << P)) {
            const float4 cs = ldg4(cf + 4 * P);  // (c, s, kappa, -)
            if constexpr (BOTH) {
              if (op.aux0 & (1 << P)) {
                const int slot = P < 4 ? ((op.aux1 >> (8 * P)) & 0xff) : op.p1();
                scratch[slot * nthr + tid] = cs.z * im_bxa<K, P>(a, b);
              }
            }
            if (op.p0() & (1 << P)) {
              xrot<K, P>(a, cs.x, cs.y);
              if constexpr (BOTH) xrot<K, P>(b, cs.x, cs.y);
            }
          }
        });
      } break;
      case OP_XROTF: {
        const bool fast = cf[3] != 0.f;
        for_each_pos<K>([&](auto pc) {
          constexpr int P = decltype(pc)::value;
          const float4 cs = ldg4(cf + 4 * P);  // identity at inactive positions
          if constexpr (BOTH) {
            if (op.aux0 & (1 << P)) {
              const int slot = P < 4 ? ((op.aux1 >> (8 * P)) & 0xff) : op.p1();
              scratch[slot * nthr + tid] = cs.z * im_bxa<K, P>(a, b);
            }
          }
          if (P < K - 1 && fast) {
            xrot_fast<K, P>(a, cs.x);
            if constexpr (BOTH) xrot_fast<K, P>(b, cs.x);
          } else {
            xrot<K, P>(a, cs.x, cs.y);
            if constexpr (BOTH) xrot<K, P>(b, cs.x, cs.y);
          }
        });
      } break;
      case OP_YROTM: if constexpr (GEN) {
        for_each_pos<K>([&](auto pc) {
          constexpr int P = decltype(pc)::value;
          if ((op.p0() | (BOTH ? op.aux0 : 0)) & (1 << P)) {
            const float4 cs = ldg4(cf + 4 * P);
            if constexpr (BOTH) {
              if (op.aux0 & (1 << P)) {
                const int slot = P < 4 ? ((op.aux1 >> (8 * P)) & 0xff) : op.p1();
                scratch[slot * nthr + tid] = cs.z * im_bya<K, P>(a, b);
              }
            }
            if (op.p0() & (1 << P)) {
              yrot<K, P>(a, cs.x, cs.y);
              if constexpr (BOTH) yrot<K, P>(b, cs.x, cs.y);
            }
          }
        });
      } break;
      case OP_MAT1: if constexpr (GEN) {
        const float4 m0 = ldg4(cf), m1 = ldg4(cf + 4);
        dispatch_pos<K>(op.p0(), [&](auto pc) {
          constexpr int P = decltype(pc)::value;
          mat1<K, P>(a, m0, m1);
          if constexpr (BOTH) mat1<K, P>(b, m0, m1);
        });
      } break;
      case OP_MAT2: if constexpr (GEN) {
        if (op.p0() == 0) {
          mat2<K, 0, false>(a, a, cf);
          if constexpr (BOTH) mat2<K, 0, false>(b, b, cf);
        } else {
          mat2<K, 2, false>(a, a, cf);
          if constexpr (BOTH) mat2<K, 2, false>(b, b, cf);
        }
      } break;
      case OP_DCONST_TAB: {
        const uint32_t idx = (gbase >> op.aux0) & (uint32_t)op.aux1;
        F = cmul(F, ldg2(cf + 2 * idx));
      } break;
      case OP_DCONST_PAIR: {
        int sel = (gbase >> op.aux0) & 1;
        if (op.aux1 >= 0) sel = 2 * sel + ((gbase >> op.aux1) & 1);
        F = cmul(F, ldg2(cf + 2 * sel));
      } break;
      case OP_DREG_TAB: {
        // table entries are (re, im, -im, im): the last pair is the packed-multiply form of the constant
        const bool pending = op.aux0 != 0;  // uniform: some thread-constant factor was folded into F
        const CK kf = make_ck(F);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 t = ldg4(cf + 4 * r);
          CK k;
          k.re = t.x;
          k.q = make_float2(t.z, t.w);
          a[r] = cmulk(a[r], k);
          if constexpr (BOTH) b[r] = cmulk(b[r], k);
          if (pending) {
            a[r] = cmulk(a[r], kf);
            if constexpr (BOTH) b[r] = cmulk(b[r], kf);
          }
        }
        F = make_float2(1.f, 0.f);
      } break;
      case OP_DAPPLY: {
        const CK kf = make_ck(F);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          a[r] = cmulk(a[r], kf);
          if constexpr (BOTH) b[r] = cmulk(b[r], kf);
        }
        F = make_float2(1.f, 0.f);
      } break;
      case OP_DCROSS: {
        const int cb = (gbase >> op.aux0) & 1;
        const float4 e = ldg4(cf + 4 * cb);
        const float2 e0 = make_float2(e.x, e.y), e1 = make_float2(e.z, e.w);
        const bool id0 = e.x == 1.f && e.y == 0.f, id1 = e.z == 1.f && e.w == 0.f;
        if (!(id0 && id1)) {
          dispatch_pos<K>(op.p0(), [&](auto pc) {
            constexpr int P = decltype(pc)::value;
            mul_sel<K, P>(a, e0, e1, id0, id1);
            if constexpr (BOTH) mul_sel<K, P>(b, e0, e1, id0, id1);
          });
        }
      } break;
      default:
        if constexpr (BOTH) {
          if (GEN && op.type() == OP_GRAD_MAT1) {
            const float4 m0 = ldg4(cf), m1 = ldg4(cf + 4);
            float v = 0.f;
            dispatch_pos<K>(op.p0(), [&](auto pc) { v = grad_mat1<K, decltype(pc)::value>(a, b, m0, m1); });
            scratch[op.gslot() * nthr + tid] = v;
          } else if (GEN && op.type() == OP_GRAD_MAT2) {
            const float g = op.p0() == 0 ? mat2<K, 0, true>(a, b, cf) : mat2<K, 2, true>(a, b, cf);
            scratch[op.gslot() * nthr + tid] = g;
          } else if (op.type() == OP_GD_BEGIN) {
            store_marginals<K>(a, b, (uint32_t)op.coef, op.gslot(), scratch, tid, nthr);
          }
        }
        break;
    }
    oi += step;
  }

  bool flush = false;
  if constexpr (BOTH) flush = ps->gd_flush_end > ps->gd_flush_begin;
  if (ngrad > 0) {
    __syncthreads();  // every thread's gradient values of this pass are in the scratch
    // Reduction tasks, two per warp at a time (independent load/add chains and interleaved shuffle trees).
    const uint32_t w = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    auto partial = [&](const OpRec& t) -> float2 {  // this lane's share of task t
      const float* src = scratch + (size_t)t.p0() * nthr;
      if (t.type() == OP_TASK_F) {
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
        for (uint32_t i = lane; i < nthr; i += 64) {
          acc0 += src[i];
          if (i + 32 < nthr) acc1 += src[i + 32];
        }
        return make_float2(acc0 + acc1, 0.f);
      }
      // complex vector, restricted to the threads whose index has every bit of the mask set: blocks of
      // 32 threads are skipped as a whole, the lane bits of the mask zero this lane's share at the end
      const float2* v = reinterpret_cast<const float2*>(src) + lane;
      const uint32_t m = (uint32_t)t.aux0, mh = m & ~31u;
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll 4
      for (uint32_t i0 = 0; i0 < nthr; i0 += 64) {
        if ((i0 & mh) == mh) acc0 = add2(acc0, v[i0]);
        if (i0 + 32 < nthr && ((i0 + 32) & mh) == mh) acc1 = add2(acc1, v[i0 + 32]);
      }
      const float2 acc = add2(acc0, acc1);
      return ((lane & m) == (m & 31u)) ? acc : make_float2(0.f, 0.f);
    };
    for (int t0 = op_end + (int)w; t0 < pv.task_end; t0 += 2 * (int)nw) {
      const int t1 = t0 + (int)nw;
      const bool two = t1 < pv.task_end;
      const OpRec ta = load_op(ops_base + t0);
      const OpRec tb = load_op(ops_base + (two ? t1 : t0));
      float2 sa = partial(ta), sb = partial(tb);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sa.x += __shfl_xor_sync(0xffffffffu, sa.x, o);
        sb.x += __shfl_xor_sync(0xffffffffu, sb.x, o);
        sa.y += __shfl_xor_sync(0xffffffffu, sa.y, o);
        sb.y += __shfl_xor_sync(0xffffffffu, sb.y, o);
      }
      if (lane == 0) {  // every task owns its slot(s) of the flush window
        cx.gacc[ta.coef] = sa.x;
        if (ta.type() == OP_TASK_C) cx.gacc[ta.coef + 1] = sa.y;
        if (two) {
          cx.gacc[tb.coef] = sb.x;
          if (tb.type() == OP_TASK_C) cx.gacc[tb.coef + 1] = sb.y;
        }
      }
    }
  }
  if (ngrad > 0 || flush) __syncthreads();  // the scratch is the tile: reductions finish before amplitudes are stored
  if constexpr (BOTH) {
    if (flush) {
      // End of a flush window: one thread per gradient slot combines the reduced sums (DevGradDesc) and
      // adds the CTA's share to the float64 accumulators.
      double* grow = ka.gacc + (size_t)(ka.per_state ? (ka.grow0 + (int)u) : 0) * ka.P;
      const float* G = cx.gacc;
      for (int g = ps->gd_flush_begin + (int)tid; g < ps->gd_flush_end; g += (int)nthr) {
        const int4 d0 = __ldg(reinterpret_cast<const int4*>(ka.gdescs + g));
        const int4 d1 = __ldg(reinterpret_cast<const int4*>(ka.gdescs + g) + 1);
        const int kind = d0.x, i_tot = d0.w & 0xffff, i_a = (int)((uint32_t)d0.w >> 16);
        float val;
        if (kind == 0) {
          val = G[i_tot];
        } else {
          const int ca = (int)(int8_t)(d1.y & 0xff), cb = (int)(int8_t)((d1.y >> 8) & 0xff);
          const bool ok_a = ca < 0 || ((goff >> ca) & 1u), ok_b = cb < 0 || ((goff >> cb) & 1u);
          const float2 tot = make_float2(G[i_tot], G[i_tot + 1]);
          const float2 A = ok_a ? make_float2(G[i_a], G[i_a + 1]) : make_float2(0.f, 0.f);
          const float* m = ka.coef + (size_t)u * ka.coef_stride + d0.z;
          const float4 m01 = __ldg(reinterpret_cast<const float4*>(m));
          if (kind == 1) {
            const float2 U0 = make_float2(tot.x - A.x, tot.y - A.y);
            val = 2.f * (m01.x * U0.x + m01.y * U0.y + m01.z * A.x + m01.w * A.y);
          } else {
            const int i_b = d1.x & 0xffff, i_ab = (int)((uint32_t)d1.x >> 16);
            const float4 m23 = __ldg(reinterpret_cast<const float4*>(m) + 1);
            const float2 B = ok_b ? make_float2(G[i_b], G[i_b + 1]) : make_float2(0.f, 0.f);
            const float2 AB = (ok_a && ok_b) ? make_float2(G[i_ab], G[i_ab + 1]) : make_float2(0.f, 0.f);
            const float2 U10 = make_float2(A.x - AB.x, A.y - AB.y), U01 = make_float2(B.x - AB.x, B.y - AB.y);
            const float2 U00 = make_float2(tot.x - A.x - U01.x, tot.y - A.y - U01.y);
            val = 2.f * (m01.x * U00.x + m01.y * U00.y + m01.z * U01.x + m01.w * U01.y +
                         m23.x * U10.x + m23.y * U10.y + m23.z * AB.x + m23.w * AB.y);
          }
        }
        if (val != 0.f) atomicAdd(grow + d0.y, (double)val);
      }
    }
  }
  // (after the last gradient pass of a launch that stores nothing, no one reads the tiles again)
  if (!(BOTH && last && !(ka.L.flags & (LF_STORE_PSI | LF_STORE_LAM)))) {
    const uint4* ep = reinterpret_cast<const uint4*>(ps->eoff8);
#pragma unroll
    for (int i = 0; i < R / 4; ++i) {
      const uint4 w = ep[i];
      const uint32_t eo[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        *reinterpret_cast<float2*>(reinterpret_cast<char*>(s_psi) + (B8 ^ eo[r])) = a[4 * i + r];
        if constexpr (BOTH) *reinterpret_cast<float2*>(reinterpret_cast<char*>(s_lam) + (B8 ^ eo[r])) = b[4 * i + r];
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// Observable pass: the Pauli strings whose X/Y part acts inside the pass's register qubits are applied
// like gates -- the partner amplitude psi[i ^ x] is another register of the same thread and the sign
// pattern over the register index is a host-built coefficient table.
// BOTH = true (adjoint kernel): h = H psi accumulates in the lambda tile (unscaled; the expectation phase
// finishes E and lambda from it).  BOTH = false: E is accumulated directly.  Single observable only.
// ---------------------------------------------------------------------------------
// MODE 0: per-amplitude coefficient table.  MODE 1: the table is uniform (e.g. a bare X or XX string: no
// Z dressing inside the registers).  MODE 2: the table is zero where the flipped register bits have even
// parity and uniform elsewhere (XX + YY with equal coefficients only connects |01> and |10>): half of the
// amplitudes are skipped.  The host sets the mode (DevOp::aux1) from the table it built.
template <int XR, bool BOTH, int MODE>
__device__ __forceinline__ void hx_apply(const float2 (&a)[16], float2 (&b)[BOTH ? 16 : 1], const float* cf,
                                         const float sgn, float& e) {
  float2 acc = make_float2(0.f, 0.f);
  if constexpr (MODE == 0) {
#pragma unroll
    for (int r0 = 0; r0 < 16; r0 += 4) {
      const float4 t4 = ldg4(cf + r0);  // coefficients are read four at a time: no table in registers
      const float tab[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + i, q = r ^ XR;
        if constexpr (BOTH) {
          b[r] = fma2(bc(sgn * tab[i]), a[q], b[r]);
        } else {
          acc = fma2(mul2(bc(tab[i]), a[r]), a[q], acc);
        }
      }
    }
    if constexpr (!BOTH) e = fmaf(sgn, hsum(acc), e);
  } else {
    constexpr int kFirst = MODE == 2 ? (XR & -XR) : 0;  // first entry of the table that carries the value
    const float t = sgn * cf[kFirst];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (MODE == 2 && !(__builtin_popcount(r & XR) & 1)) continue;
      const int q = r ^ XR;
      if constexpr (BOTH) b[r] = fma2(bc(t), a[q], b[r]);
      else acc = fma2(a[r], a[q], acc);
    }
    if constexpr (!BOTH) e = fmaf(t, hsum(acc), e);
  }
}

// The pass works on 16 amplitudes at a time (register positions 0..3 carry the flips); with K = 5 the
// fifth register position only selects the half, so the forward-only kernel never holds 32 complex
// registers here.
template <int K, bool BOTH>
__device__ __forceinline__ void run_hpass(const KernelArgs& ka, PassCtx& cx, const int p, const bool first,
                                          const bool last, const float2* s_psi, float2* s_lam, uint32_t goff,
                                          uint32_t u) {
  constexpr int R = 1 << K;
  const uint32_t tid = threadIdx.x;
  const PassView pv = begin_pass(ka, cx, p, first, last, u * ka.coef_stride);
  const DevPass* ps = pv.ps;
  const int op_begin = pv.op_begin, op_end = pv.op_end;
  const PackedOp* ops_base = pv.ops;
  const float* coef_base = pv.coef;
  uint32_t base = tid;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int sp = ps->sorted[j];
    base = ((base >> sp) << (sp + 1)) | (base & ((1u << sp) - 1u));
  }
  const uint32_t B = swz(base);
  const uint32_t gbase = goff | scatter_bits(base, ka.L.runs, ka.L.n_runs);
  float e = 0.f;
#pragma unroll 1
  for (int half = 0; half < R / 16; ++half) {
    const uint4* ep = reinterpret_cast<const uint4*>(ps->eoff8) + 4 * half;
    const uint32_t B8 = 8u * B;
    float2 a[16];
    float2 b[BOTH ? 16 : 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 w = ep[i];
      const uint32_t eo[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        a[4 * i + r] = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(s_psi) + (B8 ^ eo[r]));
        if constexpr (BOTH) {  // the first observable pass starts h = H psi from zero (the lambda tile is not cleared)
          b[4 * i + r] = first ? make_float2(0.f, 0.f)
                               : *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(s_lam) + (B8 ^ eo[r]));
        }
      }
    }
    float dsum = 0.f;
    bool any_dsum = false;  // uniform
    for (int oi = op_begin; oi < op_end; ++oi) {
      const OpRec op = load_op(ops_base + oi);
      const float* cf = coef_base + op.coef + 16 * half;
      const float sgn = (__popc(gbase & (uint32_t)op.aux0) & 1) ? -1.f : 1.f;
      const int mode = op.aux1;
      if (op.type() == OP_HD && mode == 1) {
        // a Z string that misses the register qubits: one scalar per thread; all of them are applied
        // together after the loop (one multiply-add per amplitude for the whole set)
        dsum = fmaf(sgn, cf[0], dsum);
        any_dsum = true;
        continue;
      }
      auto apply = [&](auto xc) {
        constexpr int XR = decltype(xc)::value;
        if (mode == 1) hx_apply<XR, BOTH, 1>(a, b, cf, sgn, e);
        else if (mode == 2 && XR != 0) hx_apply<XR, BOTH, (XR != 0 ? 2 : 0)>(a, b, cf, sgn, e);
        else hx_apply<XR, BOTH, 0>(a, b, cf, sgn, e);
      };
      // register xor mask: 0 for diagonal strings (OP_HD), else one or two of the low four bits
      switch (op.type() == OP_HD ? 0 : op.p0()) {
        case 0: apply(IntC<0>{}); break;
        case 1: apply(IntC<1>{}); break;
        case 2: apply(IntC<2>{}); break;
        case 3: apply(IntC<3>{}); break;
        case 4: apply(IntC<4>{}); break;
        case 5: apply(IntC<5>{}); break;
        case 6: apply(IntC<6>{}); break;
        case 8: apply(IntC<8>{}); break;
        case 9: apply(IntC<9>{}); break;
        case 10: apply(IntC<10>{}); break;
        case 12: apply(IntC<12>{}); break;
        default: break;
      }
    }
    if (any_dsum) {
      if constexpr (BOTH) {
#pragma unroll
        for (int r = 0; r < 16; ++r) b[r] = fma2(bc(dsum), a[r], b[r]);
      } else {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 16; ++r) acc = fma2(a[r], a[r], acc);
        e = fmaf(dsum, hsum(acc), e);
      }
    }
    if constexpr (BOTH) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 w = ep[i];  // re-read: 16 offsets must not stay live across the op loop
        const uint32_t eo[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
          *reinterpret_cast<float2*>(reinterpret_cast<char*>(s_lam) + (B8 ^ eo[r])) = b[4 * i + r];
      }
    }
  }
  if constexpr (!BOTH) {
    e = warp_sum(e);
    if ((tid & 31) == 0) atomicAdd(&ka.eacc[(size_t)u * ka.O], (double)e);
  }
}

__device__ __forceinline__ uint32_t wsw(uint32_t i) { return i ^ ((i >> 5) & 31u); }  // float scratch swizzle

// In-place Walsh-Hadamard transform of w[2^T] (swizzled with wsw) by the whole CTA, K bits per pass.
template <int K>
__device__ __forceinline__ void wht_inplace(float* w, const int T, const uint32_t tid, const uint32_t nthr) {
  for (int lo = 0; lo < T; lo += K) {
    const int kb = (T - lo) < K ? (T - lo) : K;
    const uint32_t ngroups = 1u << (T - kb);
    for (uint32_t g = tid; g < ngroups; g += nthr) {
      const uint32_t base = ((g >> lo) << (lo + kb)) | (g & ((1u << lo) - 1u));
      float v[1 << K];
#pragma unroll
      for (int r = 0; r < (1 << K); ++r)
        if (r < (1 << kb)) v[r] = w[wsw(base | ((uint32_t)r << lo))];
#pragma unroll
      for (int st = 0; st < K; ++st) {
        if (st < kb) {
#pragma unroll
          for (int r = 0; r < (1 << K); ++r) {
            if (!(r & (1 << st)) && (r | (1 << st)) < (1 << kb)) {
              const float x = v[r], y = v[r | (1 << st)];
              v[r] = x + y;
              v[r | (1 << st)] = x - y;
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < (1 << K); ++r)
        if (r < (1 << kb)) w[wsw(base | ((uint32_t)r << lo))] = v[r];
    }
    __syncthreads();
  }
}

// All diagonal (Z-string) terms at once: W = WHT(|psi|^2) gives sum_i (-1)^{parity(i & z)} |psi_i|^2 for
// every in-tile mask z; the adjoint factor D_i = sum_terms g c (-1)^{parity(i & z)} is the WHT of the
// sparse vector V[z] = g c.  Out-of-tile bits of z only contribute a per-tile sign.
template <int K, bool ADJ>
__device__ __forceinline__ void wht_diag(const KernelArgs& ka, const float2* s_psi, float* W, float* V,
                                         float (&dgall)[ADJ ? (1 << K) : 1], const uint32_t goff, const uint32_t u,
                                         const bool want_lam) {
  constexpr int R = 1 << K;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  const uint32_t ph_tid = swz(tid);
  const uint32_t tmask = (1u << ka.T) - 1u;
#pragma unroll
  for (int m = 0; m < R; ++m) {
    const float2 a = s_psi[ph_tid ^ ka.L.soff[m]];
    const uint32_t i = wsw((uint32_t)m * nthr | tid);
    W[i] = a.x * a.x + a.y * a.y;
    if constexpr (ADJ) V[i] = 0.f;
  }
  __syncthreads();
  wht_inplace<K>(W, ka.T, tid, nthr);
  for (int t = (int)tid; t < ka.n_dterms; t += (int)nthr) {
    const int4 d = __ldg(reinterpret_cast<const int4*>(ka.dterms + t));
    const uint32_t z = (uint32_t)d.y;
    float val = __int_as_float(d.x);
    if (__popc(goff & z) & 1) val = -val;
    const uint32_t zi = wsw(z & tmask);
    atomicAdd(&ka.eacc[(size_t)u * ka.O + d.z], (double)(val * W[zi]));
    if constexpr (ADJ) {
      if (want_lam) atomicAdd(&V[zi], __ldg(&ka.dgrad[(size_t)u * ka.O + d.z]) * val);
    }
  }
  __syncthreads();
  if constexpr (ADJ) {
    wht_inplace<K>(V, ka.T, tid, nthr);
#pragma unroll
    for (int m = 0; m < R; ++m) dgall[m] = V[wsw((uint32_t)m * nthr | tid)];
    __syncthreads();  // V is about to be overwritten by the lambda tile
  }
}

// Coefficients c(i_m) = k0 + sum_t k_t (-1)^{parity(i_m & z_t)} of one x-group for MC amplitudes.
// A term whose z-mask does not touch the bits that distinguish the thread's amplitudes has the same
// sign for all of them (warp-uniform test on the host-built sign word): it is added to a scalar.
template <int MC, bool CPLX>
__device__ __forceinline__ void group_coefficients(const DevTerm* terms, float (&cr)[MC], float (&ci)[CPLX ? MC : 1],
                                                   const uint32_t gi_tid, const int m0, const int t0, const int t1,
                                                   const float k0r, const float k0i) {
  constexpr uint32_t kAll = MC >= 32 ? 0xffffffffu : ((1u << MC) - 1u);
  float c0r = k0r, c0i = k0i;
#pragma unroll
  for (int m = 0; m < MC; ++m) {
    cr[m] = 0.f;
    if constexpr (CPLX) ci[m] = 0.f;
  }
  for (int t = t0; t < t1; ++t) {
    const float4 tv = *reinterpret_cast<const float4*>(terms + t);
    const uint32_t tp = (uint32_t)(__popc(gi_tid & __float_as_uint(tv.z)) & 1) << 31;
    const uint32_t word = (__float_as_uint(tv.w) >> m0) & kAll;
    // thread parity folded into k once per term; the per-amplitude bit then only selects +k or -k
    const float kx = __uint_as_float(__float_as_uint(tv.x) ^ tp), ky = __uint_as_float(__float_as_uint(tv.y) ^ tp);
    if (word == 0u || word == kAll) {
      c0r += word ? -kx : kx;
      if constexpr (CPLX) c0i += word ? -ky : ky;
      continue;
    }
#pragma unroll
    for (int m = 0; m < MC; ++m) {
      const bool neg = (word >> m) & 1u;
      cr[m] += neg ? -kx : kx;
      if constexpr (CPLX) ci[m] += neg ? -ky : ky;
    }
  }
#pragma unroll
  for (int m = 0; m < MC; ++m) {
    cr[m] += c0r;
    if constexpr (CPLX) ci[m] += c0i;
  }
}

// h[m] += c(i_m) * psi[i_m ^ x] for one off-diagonal x-group.  GLOBAL: the partner lives in
// another tile (read through L2), else inside this tile's shared memory.
// True when no term of the group distinguishes the thread's MC amplitudes (warp-uniform): the
// coefficient is then one scalar per thread, returned in (c0r, c0i).
template <int MC, bool CPLX>
__device__ __forceinline__ bool uniform_coefficient(const DevTerm* terms, const uint32_t gi_tid, const int m0,
                                                    const int t0, const int t1, float& c0r, float& c0i) {
  constexpr uint32_t kAll = MC >= 32 ? 0xffffffffu : ((1u << MC) - 1u);
  for (int t = t0; t < t1; ++t) {
    const float4 tv = *reinterpret_cast<const float4*>(terms + t);
    const uint32_t word = (__float_as_uint(tv.w) >> m0) & kAll;
    if (word != 0u && word != kAll) return false;
    const uint32_t sg = (uint32_t)((__popc(gi_tid & __float_as_uint(tv.z)) + (int)(word & 1u)) & 1) << 31;
    c0r += __uint_as_float(__float_as_uint(tv.x) ^ sg);
    if constexpr (CPLX) c0i += __uint_as_float(__float_as_uint(tv.y) ^ sg);
  }
  return true;
}

// Scalar-coefficient fast path of an x-group (no term distinguishes the thread's amplitudes): also in the
// adjoint kernel since the groups that reach it are the out-of-tile ones, where a zero scalar skips the
// cross-tile reads (measured 15.01 -> 14.89 ms per 4096 bitstrings on config 3, gpurun_out/s6).
#ifndef QHBM_ADJ_SCALAR
#define QHBM_ADJ_SCALAR 1
#endif
template <bool ADJ>
constexpr bool kScalarGroups = !ADJ || QHBM_ADJ_SCALAR;


template <int MC, bool CPLX, bool GLOBAL, bool SCALAR>
__device__ __forceinline__ void group_offdiag(const KernelArgs& ka, const DevTerm* terms, const float2* s_psi,
                                              const float2* __restrict__ psi_u, float2 (&h)[MC],
                                              const uint32_t gi_tid, const uint32_t ph_tid,
                                              const uint32_t nthr, const int m0, const uint32_t x, const int xl,
                                              const int t0, const int t1, const float k0r, const float k0i) {
  const uint32_t pxor = GLOBAL ? 0u : swz((uint32_t)xl);
  if constexpr (SCALAR) {  // forward-only kernel: in the adjoint kernel the extra code costs more than it saves
    float c0r = k0r, c0i = k0i;
    if (uniform_coefficient<MC, CPLX>(terms, gi_tid, m0, t0, t1, c0r, c0i)) {
      if (c0r == 0.f && (!CPLX || c0i == 0.f)) return;  // e.g. XX + YY on an aligned pair: nothing to add
#pragma unroll
      for (int m = 0; m < MC; ++m) {
        float2 p;
        if constexpr (GLOBAL) p = psi_u[(gi_tid | ka.L.moff[m0 + m]) ^ x];
        else p = s_psi[ph_tid ^ ka.L.soff[m0 + m] ^ pxor];
        h[m] = fma2(bc(c0r), p, h[m]);
        if constexpr (CPLX) h[m] = fma2(np_pair(c0i), swp(p), h[m]);
      }
      return;
    }
  }
  float cr[MC];
  float ci[CPLX ? MC : 1];
  group_coefficients<MC, CPLX>(terms, cr, ci, gi_tid, m0, t0, t1, k0r, k0i);
#pragma unroll
  for (int m = 0; m < MC; ++m) {
    float2 p;
    if constexpr (GLOBAL) p = psi_u[(gi_tid | ka.L.moff[m0 + m]) ^ x];
    else p = s_psi[ph_tid ^ ka.L.soff[m0 + m] ^ pxor];
    h[m] = fma2(bc(cr[m]), p, h[m]);
    if constexpr (CPLX) h[m] = fma2(np_pair(ci[m]), swp(p), h[m]);
  }
}

// ---------------------------------------------------------------------------------
// Expectation phase: E_j = Re <psi|H_j|psi>, and (adjoint) lambda = sum_j g_j H_j psi.
// H psi[i] = sum_groups c_g(i) psi[i ^ x_g]; all terms of a group share the x-mask.
// The thread's amplitudes are i_m = m * nthreads + tid; parity(i_m & z) splits into a per-thread
// bit and a per-m bit that the host precomputed (DevTerm::mword).  Diagonal groups (x = 0) only
// need |psi_i|^2 and a real per-amplitude factor D_i = sum_j g_j c_j(i), applied once at the end.
// ---------------------------------------------------------------------------------
template <int K, bool ADJ>
__device__ __forceinline__ void expect_phase(const KernelArgs& ka, float2* s_psi, float2* s_lam, float4* s_stage,
                                             uint32_t goff, uint32_t u, const float2* __restrict__ psi_u,
                                             const bool hinit) {
  constexpr int R = 1 << K;
  constexpr int MC = ADJ ? (R < 8 ? R : 8) : (R < 16 ? R : 16);  // amplitudes per thread handled at a time
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  const bool want_lam = ADJ && ka.dgrad != nullptr;
  const uint32_t gi_tid = goff | scatter_bits(tid, ka.L.runs, ka.L.n_runs);
  const uint32_t ph_tid = swz(tid);
  __syncthreads();
  // stage this launch's slice of the Pauli tables (group headers: 2 float4 each, terms: 1 float4 each)
  const DevTermGroup* groups = ka.groups;
  const DevTerm* terms = ka.terms;
  {
    const int ng = ka.L.grp_end - ka.L.grp_begin, nt = ka.L.term_end - ka.L.term_begin;
    if (2 * ng + nt <= kStageOps + kStageCoef / 4) {
      const float4* gg = reinterpret_cast<const float4*>(ka.groups + ka.L.grp_begin);
      const float4* gt = reinterpret_cast<const float4*>(ka.terms + ka.L.term_begin);
      for (int i = (int)tid; i < 2 * ng; i += (int)nthr) s_stage[i] = __ldg(gg + i);
      for (int i = (int)tid; i < nt; i += (int)nthr) s_stage[2 * ng + i] = __ldg(gt + i);
      groups = reinterpret_cast<const DevTermGroup*>(s_stage) - ka.L.grp_begin;
      terms = reinterpret_cast<const DevTerm*>(s_stage + 2 * ng) - ka.L.term_begin;
      __syncthreads();
    }
  }
  float dgall[ADJ ? R : 1];
  const bool wht = ka.n_dterms > 0 && ka.L.expect_stage == 0;  // diagonal terms live in stage 0
  if (wht) {
    // float scratch: the (not yet written) lambda tile, or the extra tile of the forward-only kernel
    float* W = reinterpret_cast<float*>(s_psi + (1u << ka.T));
    wht_diag<K, ADJ>(ka, s_psi, W, W + (1u << ka.T), dgall, goff, u, want_lam);
  }
#pragma unroll
  for (int m0 = 0; m0 < R; m0 += MC) {
    float2 a[MC];
    float p2[MC];
    float2 lam[ADJ ? MC : 1];
    float dg[ADJ ? MC : 1];
    float p2sum = 0.f;
#pragma unroll
    for (int m = 0; m < MC; ++m) {
      a[m] = s_psi[ph_tid ^ ka.L.soff[m0 + m]];
      p2[m] = a[m].x * a[m].x + a[m].y * a[m].y;
      if constexpr (!ADJ) p2sum += p2[m];
      if constexpr (ADJ) {
        lam[m] = make_float2(0.f, 0.f);
        dg[m] = wht ? dgall[m0 + m] : 0.f;  // m0 is a constant after unrolling the chunk loop
      }
    }
    for (int ri = ka.L.rng_begin; ri < ka.L.rng_end; ++ri) {
      const int4 orng = __ldg(reinterpret_cast<const int4*>(ka.opranges + ri));
      const int j = orng.z;
      const float gj = want_lam ? __ldg(&ka.dgrad[(size_t)u * ka.O + j]) : 0.f;
      float ej = 0.f;
      float2 ej2 = make_float2(0.f, 0.f);
      float2 h[MC];
      bool offdiag = false;
      if constexpr (ADJ) {
        if (hinit) {  // the observable passes left their part of H psi in the lambda tile (single observable)
          offdiag = true;
#pragma unroll
          for (int m = 0; m < MC; ++m) h[m] = s_lam[ph_tid ^ ka.L.soff[m0 + m]];
        }
      }
      const int g_end = orng.y;
      int g = orng.x;
      int4 gh, gk;
      if (g < g_end) {
        gh = reinterpret_cast<const int4*>(groups + g)[0];
        gk = reinterpret_cast<const int4*>(groups + g)[1];
      }
      for (; g < g_end; ++g) {
        const int4 ch = gh, ck = gk;
        if (g + 1 < g_end) {  // prefetch the next group header
          gh = reinterpret_cast<const int4*>(groups + g + 1)[0];
          gk = reinterpret_cast<const int4*>(groups + g + 1)[1];
        }
        const uint32_t x = (uint32_t)ch.x;
        const int xl = ch.y;
        const float k0r = __int_as_float(ck.x), k0i = __int_as_float(ck.y);
        if (x == 0) {
          if constexpr (!ADJ) {
            float c0r = k0r, c0i = 0.f;
            if (uniform_coefficient<MC, false>(terms, gi_tid, m0, ch.z, ch.w, c0r, c0i)) {
              ej = fmaf(c0r, p2sum, ej);
              continue;
            }
          }
          float cr[MC], ci[1];
          group_coefficients<MC, false>(terms, cr, ci, gi_tid, m0, ch.z, ch.w, k0r, 0.f);
#pragma unroll
          for (int m = 0; m < MC; ++m) {
            ej = fmaf(cr[m], p2[m], ej);
            if constexpr (ADJ) dg[m] = fmaf(gj, cr[m], dg[m]);
          }
          continue;
        }
        if (!offdiag) {
          offdiag = true;
#pragma unroll
          for (int m = 0; m < MC; ++m) h[m] = make_float2(0.f, 0.f);
        }
        if (ck.z == 0) {
          if (xl >= 0) group_offdiag<MC, false, false, kScalarGroups<ADJ>>(ka, terms, s_psi, psi_u, h, gi_tid, ph_tid, nthr, m0, x, xl, ch.z, ch.w, k0r, k0i);
          else group_offdiag<MC, false, true, kScalarGroups<ADJ>>(ka, terms, s_psi, psi_u, h, gi_tid, ph_tid, nthr, m0, x, xl, ch.z, ch.w, k0r, k0i);
        } else {
          if (xl >= 0) group_offdiag<MC, true, false, kScalarGroups<ADJ>>(ka, terms, s_psi, psi_u, h, gi_tid, ph_tid, nthr, m0, x, xl, ch.z, ch.w, k0r, k0i);
          else group_offdiag<MC, true, true, kScalarGroups<ADJ>>(ka, terms, s_psi, psi_u, h, gi_tid, ph_tid, nthr, m0, x, xl, ch.z, ch.w, k0r, k0i);
        }
      }
      if (offdiag) {
#pragma unroll
        for (int m = 0; m < MC; ++m) {
          ej2 = fma2(a[m], h[m], ej2);
          if constexpr (ADJ) lam[m] = fma2(bc(gj), h[m], lam[m]);
        }
      }
      ej = warp_sum(ej + hsum(ej2));
      if ((tid & 31) == 0) atomicAdd(&ka.eacc[(size_t)u * ka.O + j], (double)ej);
    }
    if constexpr (ADJ) {
#pragma unroll
      for (int m = 0; m < MC; ++m) {
        lam[m] = fma2(bc(dg[m]), a[m], lam[m]);
        s_lam[ph_tid ^ ka.L.soff[m0 + m]] = lam[m];
      }
    }
  }
}

// `sparse` != 0 (LF_SPARSE_IN, psi only): amplitudes whose index differs from `basis` in those bits are known
// to be zero and were never stored: they are zero-filled without touching memory.
template <int K>
__device__ __forceinline__ void load_tile(float2* s, const float2* __restrict__ g, uint32_t goff, const KernelArgs& ka,
                                          const uint32_t sparse = 0u, const uint32_t basis = 0u) {
  constexpr int R = 1 << K;
  const uint32_t tid = threadIdx.x;
  const uint32_t gt = goff | scatter_bits(tid, ka.L.runs, ka.L.n_runs);
  const uint32_t pt = swz(tid);
  if (sparse) {
    if (ka.async_tile) {
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const uint32_t gi = gt | ka.L.moff[m];
        __pipeline_memcpy_async(s + (pt ^ ka.L.soff[m]), g + gi, 8, ((gi ^ basis) & sparse) ? 8 : 0);
      }
      __pipeline_commit();
    } else {
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const uint32_t gi = gt | ka.L.moff[m];
        s[pt ^ ka.L.soff[m]] = ((gi ^ basis) & sparse) ? make_float2(0.f, 0.f) : g[gi];
      }
    }
    return;
  }
  if (ka.async_tile) {
    // Measured in profiles/r2_tile_copy_experiment.md: global -> shared without the register round trip,
    // 8-byte cp.async per amplitude (the nibble-XOR swizzle permutes amplitudes inside 128-byte groups, so a
    // larger copy unit -- 16-byte cp.async or a TMA box -- would land in the wrong order).  The copies
    // complete at the first pass's program wait (cp.async.wait_all) + barrier.
#pragma unroll
    for (int m = 0; m < R; ++m) __pipeline_memcpy_async(s + (pt ^ ka.L.soff[m]), g + (gt | ka.L.moff[m]), 8);
    __pipeline_commit();
    return;
  }
  float2 v[R];
#pragma unroll
  for (int m = 0; m < R; ++m) v[m] = g[gt | ka.L.moff[m]];
#pragma unroll
  for (int m = 0; m < R; ++m) s[pt ^ ka.L.soff[m]] = v[m];
}
template <int K>
__device__ __forceinline__ void store_tile(const float2* s, float2* __restrict__ g, uint32_t goff, const KernelArgs& ka,
                                           const float2 scale) {
  constexpr int R = 1 << K;
  const uint32_t tid = threadIdx.x;
  const uint32_t gt = goff | scatter_bits(tid, ka.L.runs, ka.L.n_runs);
  const uint32_t pt = swz(tid);
  const bool scaled = scale.x != 1.f || scale.y != 0.f;
#pragma unroll
  for (int m = 0; m < R; ++m) {
    float2 v = s[pt ^ ka.L.soff[m]];
    if (scaled) v = cmul(v, scale);
    g[gt | ka.L.moff[m]] = v;
  }
}

template <int K, bool ADJ>
constexpr int sweep_max_threads() { return ADJ ? (1 << (13 - K)) : 512; }

// grid = chunk * 2^(n-T) CTAs, block = 2^(T-K) threads, dynamic smem = (ADJ ? 2 : 1) * 8 * 2^T bytes.
// DENSE (forward kernel, K = 4, <= 256 threads): compiled for three resident CTAs per SM with two program
// buffers; it runs the forward sweeps of ADJOINT plans, which hold one 32 KiB psi tile each and would
// otherwise occupy the SM with the two-CTA, 128-register adjoint kernel.
template <int K, bool ADJ, bool DENSE = false, bool GEN = true>
__global__ void __launch_bounds__(DENSE ? 256 : sweep_max_threads<K, ADJ>(), DENSE ? 3 : 1)
    sweep_kernel(const __grid_constant__ KernelArgs ka) {
  static_assert(!(DENSE && ADJ), "the dense variant is forward-only");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kOpsCap = ADJ ? kStageOpsAdj : kStageOps;
  __shared__ float4 s_stage[((ADJ || DENSE) ? 2 : 1) * PassCtx::stage_f4(kOpsCap)];
  __shared__ float s_gacc[ADJ ? kGaccFloats : 1];
  float2* s_psi = reinterpret_cast<float2*>(smem_raw);
  float2* s_lam = s_psi + (ADJ ? (1u << ka.T) : 0u);
  constexpr int R = 1 << K;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  const uint32_t tshift = (uint32_t)(ka.n - ka.T);
  // LF_SPARSE_OUT (first forward sweep of a multi-sweep computation): one CTA per state, on the tile that
  // holds its basis index; the other tiles stay all-zero and are neither computed nor stored.
  const bool one_tile = (ka.L.flags & LF_SPARSE_OUT) != 0u;
  const uint32_t u = one_tile ? blockIdx.x : blockIdx.x >> tshift;
  const uint32_t goff = one_tile ? ((uint32_t)ka.basis[u] & ~ka.L.tile_mask)
                                 : scatter_bits(blockIdx.x & ((1u << tshift) - 1u), ka.L.oruns, ka.L.n_oruns);
  float2* psi_u = ka.psi ? ka.psi + ((size_t)u << ka.n) : nullptr;
  float2* lam_u = ka.lam ? ka.lam + ((size_t)u << ka.n) : nullptr;
  const uint32_t flags = ka.L.flags;
  const float2 one = make_float2(1.f, 0.f);
  PassCtx cx;
  cx.stage = s_stage;
  cx.gacc = s_gacc;
  cx.buf = 0;
  cx.dbuf = ADJ || DENSE;
  cx.ops_cap = kOpsCap;
  __shared__ __align__(8) uint64_t s_bar[2];
  cx.bars = s_bar;
  cx.phase = 0u;
  if (ka.bulk_stage && tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();  // (the first begin_pass starts with a barrier before any copy is issued)
  }

  bool active = true;
  if (flags & LF_INIT_BASIS) {
    const uint32_t basis = (uint32_t)ka.basis[u];
    active = (basis & ~ka.L.tile_mask) == goff;
    if (!active && (flags & LF_SPARSE_OUT)) return;  // (cannot happen: the grid only holds the active tiles)
    const uint32_t lb = gather_bits(basis, ka.L.runs, ka.L.n_runs);
    const uint32_t pt = swz(tid);
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const uint32_t l = (uint32_t)m * nthr | tid;
      s_psi[pt ^ ka.L.soff[m]] = make_float2((active && l == lb) ? 1.f : 0.f, 0.f);
    }
  } else if (flags & LF_LOAD_PSI) {
    if (flags & LF_SPARSE_IN) load_tile<K>(s_psi, psi_u, goff, ka, ka.L.sparse_mask, (uint32_t)ka.basis[u]);
    else load_tile<K>(s_psi, psi_u, goff, ka);
  }
  if constexpr (ADJ) {
    if (flags & LF_LOAD_LAM) load_tile<K>(s_lam, lam_u, goff, ka);
  }

  if (active) {
    for (int p = ka.L.pass_a_begin; p < ka.L.pass_a_end; ++p)
      run_pass<K, false, GEN>(ka, cx, p, p == ka.L.pass_a_begin, p + 1 == ka.L.pass_a_end, s_psi, s_lam, goff, u);
  }
  if (ka.async_tile) __pipeline_wait_prior(0);  // (experiment switch) tile copies that no pass has waited for
  if constexpr (!DENSE) {  // (the dense forward variant only ever runs plain forward sweeps)
  if (flags & LF_WRITE_STATE) {
    __syncthreads();
    const float2 phase = ka.phase_coef >= 0 ? ldg2(ka.coef + (size_t)u * ka.coef_stride + ka.phase_coef) : one;
    store_tile<K>(s_psi, ka.state_out + ((size_t)u << ka.n), goff, ka, phase);
  }
  if (flags & LF_EXPECT) {
    const bool hpasses = ka.L.pass_h_end > ka.L.pass_h_begin;
    if (hpasses) {
      for (int p = ka.L.pass_h_begin; p < ka.L.pass_h_end; ++p)
        run_hpass<K, ADJ>(ka, cx, p, p == ka.L.pass_h_begin, p + 1 == ka.L.pass_h_end, s_psi, s_lam, goff, u);
    }
    expect_phase<K, ADJ>(ka, s_psi, s_lam, s_stage, goff, u, psi_u, hpasses);
  }
  }
  if constexpr (ADJ) {
    for (int p = ka.L.pass_b_begin; p < ka.L.pass_b_end; ++p)
      run_pass<K, true, GEN>(ka, cx, p, p == ka.L.pass_b_begin, p + 1 == ka.L.pass_b_end, s_psi, s_lam, goff, u);
  }
  if (flags & (LF_STORE_PSI | LF_STORE_LAM)) {
    __syncthreads();
    if (flags & LF_STORE_PSI) store_tile<K>(s_psi, ka.psi_out + ((size_t)u << ka.n), goff, ka, one);
    if constexpr (ADJ) {
      if (flags & LF_STORE_LAM) store_tile<K>(s_lam, lam_u, goff, ka, one);
    }
  }
}

#ifndef QHBM_SWEEP_KERNELS_ONLY  // (sim_lean.cu only instantiates the sweep kernels)
#include "prep_kernels.cuh"
#endif  // QHBM_SWEEP_KERNELS_ONLY

}  // namespace qhbm
