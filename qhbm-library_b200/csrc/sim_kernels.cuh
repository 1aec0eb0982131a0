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
#include <cuda_runtime.h>
#include <stdint.h>

#include "gate_math.h"
#include "program.h"

namespace qhbm {

struct KernelArgs {
  LaunchDesc L;
  const DevPass* passes;
  const DevOp* ops;
  const float* coef;
  const int32_t* gsym;
  const DevTerm* terms;
  const DevTermGroup* groups;
  const DevOpRange* opranges;
  float2* psi;            // [chunk][2^n] workspace (multi-tile only)
  float2* lam;            // [chunk][2^n] workspace (multi-tile adjoint only)
  const uint64_t* basis;  // [chunk]
  const float* dgrad;     // [chunk, O] upstream gradients (adjoint) or nullptr
  double* eacc;           // [chunk, O] expectation accumulators
  double* gacc;           // [rows, P] gradient accumulators
  float2* state_out;      // debug
  int32_t n, T, O, P;
  int32_t grow0;          // accumulator row of this chunk's first state (per-state gradients)
  int32_t per_state;
};

__device__ __forceinline__ uint32_t swz(uint32_t x) {
  return (x & ~15u) | ((x ^ (x >> 4) ^ (x >> 8) ^ (x >> 12)) & 15u);
}
__device__ __forceinline__ uint32_t scatter_bits(uint32_t l, const BitRun* runs, int nr) {
  uint32_t g = 0;
  for (int i = 0; i < nr; ++i)
    g |= ((l >> runs[i].local_start) & ((1u << runs[i].len) - 1u)) << runs[i].global_start;
  return g;
}
__device__ __forceinline__ uint32_t gather_bits(uint32_t g, const BitRun* runs, int nr) {
  uint32_t l = 0;
  for (int i = 0; i < nr; ++i)
    l |= ((g >> runs[i].global_start) & ((1u << runs[i].len) - 1u)) << runs[i].local_start;
  return l;
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

// ---------------------------------------------------------------------------------
// Register-level gate blocks.  P / LO are compile-time register positions.
// ---------------------------------------------------------------------------------
template <int K, int P>
__device__ __forceinline__ void mat1(float2 (&a)[1 << K], const float4 m0, const float4 m1) {
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) {
    if (r & (1 << P)) continue;
    const float2 x0 = a[r], x1 = a[r | (1 << P)];
    a[r].x = m0.x * x0.x - m0.y * x0.y + m0.z * x1.x - m0.w * x1.y;
    a[r].y = m0.x * x0.y + m0.y * x0.x + m0.z * x1.y + m0.w * x1.x;
    a[r | (1 << P)].x = m1.x * x0.x - m1.y * x0.y + m1.z * x1.x - m1.w * x1.y;
    a[r | (1 << P)].y = m1.x * x0.y + m1.y * x0.x + m1.z * x1.y + m1.w * x1.x;
  }
}
template <int K>
__device__ __forceinline__ void mat1_dyn(float2 (&a)[1 << K], int p, const float4 m0, const float4 m1) {
  switch (p) {
    case 0: mat1<K, 0>(a, m0, m1); break;
    case 1: mat1<K, 1>(a, m0, m1); break;
    case 2: mat1<K, 2>(a, m0, m1); break;
    case 3: mat1<K, 3>(a, m0, m1); break;
    default: if constexpr (K > 4) mat1<K, 4>(a, m0, m1); break;
  }
}

// 2 Re sum_r conj(b_r) (M a)_r over the thread's amplitudes.
template <int K, int P>
__device__ __forceinline__ float grad_mat1(const float2 (&a)[1 << K], const float2 (&b)[1 << K],
                                           const float4 m0, const float4 m1) {
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) {
    if (r & (1 << P)) continue;
    const float2 x0 = a[r], x1 = a[r | (1 << P)];
    const float y0x = m0.x * x0.x - m0.y * x0.y + m0.z * x1.x - m0.w * x1.y;
    const float y0y = m0.x * x0.y + m0.y * x0.x + m0.z * x1.y + m0.w * x1.x;
    const float y1x = m1.x * x0.x - m1.y * x0.y + m1.z * x1.x - m1.w * x1.y;
    const float y1y = m1.x * x0.y + m1.y * x0.x + m1.z * x1.y + m1.w * x1.x;
    s += b[r].x * y0x + b[r].y * y0y + b[r | (1 << P)].x * y1x + b[r | (1 << P)].y * y1y;
  }
  return 2.f * s;
}
template <int K>
__device__ __forceinline__ float grad_mat1_dyn(const float2 (&a)[1 << K], const float2 (&b)[1 << K], int p,
                                               const float4 m0, const float4 m1) {
  switch (p) {
    case 0: return grad_mat1<K, 0>(a, b, m0, m1);
    case 1: return grad_mat1<K, 1>(a, b, m0, m1);
    case 2: return grad_mat1<K, 2>(a, b, m0, m1);
    case 3: return grad_mat1<K, 3>(a, b, m0, m1);
    default: if constexpr (K > 4) return grad_mat1<K, 4>(a, b, m0, m1);
  }
  return 0.f;
}

// 4x4 block on register positions (LO+1, LO); matrix index = 2*bit(LO+1) + bit(LO).
template <int K, int LO, bool GRAD>
__device__ __forceinline__ float mat2(float2 (&a)[1 << K], const float2 (&b)[1 << K], const float* __restrict__ mp) {
  float2 m[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 v = ldg4(mp + 4 * i);
    m[2 * i] = make_float2(v.x, v.y);
    m[2 * i + 1] = make_float2(v.z, v.w);
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) {
    if (r & (3 << LO)) continue;
    float2 x[4], y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = a[r | (j << LO)];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float yr = 0.f, yi = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        yr += m[4 * i + j].x * x[j].x - m[4 * i + j].y * x[j].y;
        yi += m[4 * i + j].x * x[j].y + m[4 * i + j].y * x[j].x;
      }
      y[i] = make_float2(yr, yi);
    }
    if constexpr (GRAD) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s += b[r | (i << LO)].x * y[i].x + b[r | (i << LO)].y * y[i].y;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) a[r | (i << LO)] = y[i];
    }
  }
  return 2.f * s;
}

template <int K, int P>
__device__ __forceinline__ void mul_sel(float2 (&a)[1 << K], const float2 e0, const float2 e1) {
#pragma unroll
  for (int r = 0; r < (1 << K); ++r) a[r] = cmul(a[r], (r & (1 << P)) ? e1 : e0);
}
template <int K>
__device__ __forceinline__ void mul_sel_dyn(float2 (&a)[1 << K], int p, const float2 e0, const float2 e1) {
  switch (p) {
    case 0: mul_sel<K, 0>(a, e0, e1); break;
    case 1: mul_sel<K, 1>(a, e0, e1); break;
    case 2: mul_sel<K, 2>(a, e0, e1); break;
    case 3: mul_sel<K, 3>(a, e0, e1); break;
    default: if constexpr (K > 4) mul_sel<K, 4>(a, e0, e1); break;
  }
}

template <int K, int P>
__device__ __forceinline__ float sum_bit(const float (&u)[1 << K]) {
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < (1 << K); ++r)
    if (r & (1 << P)) s += u[r];
  return s;
}
template <int K>
__device__ __forceinline__ float sum_bit_dyn(const float (&u)[1 << K], int p) {
  switch (p) {
    case 0: return sum_bit<K, 0>(u);
    case 1: return sum_bit<K, 1>(u);
    case 2: return sum_bit<K, 2>(u);
    case 3: return sum_bit<K, 3>(u);
    default: if constexpr (K > 4) return sum_bit<K, 4>(u);
  }
  return 0.f;
}
template <int K, int PH, int PL>
__device__ __forceinline__ float sum_both(const float (&u)[1 << K]) {
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < (1 << K); ++r)
    if ((r & (1 << PH)) && (r & (1 << PL))) s += u[r];
  return s;
}
template <int K>
__device__ __forceinline__ float sum_both_dyn(const float (&u)[1 << K], int ph, int pl) {
  switch (ph * 8 + pl) {
    case 8 + 0: return sum_both<K, 1, 0>(u);
    case 16 + 0: return sum_both<K, 2, 0>(u);
    case 16 + 1: return sum_both<K, 2, 1>(u);
    case 24 + 0: return sum_both<K, 3, 0>(u);
    case 24 + 1: return sum_both<K, 3, 1>(u);
    case 24 + 2: return sum_both<K, 3, 2>(u);
    default:
      if constexpr (K > 4) {
        switch (pl) {
          case 0: return sum_both<K, 4, 0>(u);
          case 1: return sum_both<K, 4, 1>(u);
          case 2: return sum_both<K, 4, 2>(u);
          default: return sum_both<K, 4, 3>(u);
        }
      }
  }
  return 0.f;
}

// Run of diagonal-gate gradient ops: they all need only w_r = conj(lam_r) psi_r, which no
// diagonal gate changes, so w is formed once per run.
template <int K>
__device__ __noinline__ void grad_diag_run(const float2 (&a)[1 << K], const float2 (&b)[1 << K],
                                           const DevOp* __restrict__ ops, int count,
                                           const float* __restrict__ coef, float* scratch, uint32_t gbase,
                                           uint32_t tid, uint32_t nthr) {
  constexpr int R = 1 << K;
  float u[R], v[R];
  float U = 0.f, V = 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    u[r] = b[r].x * a[r].x + b[r].y * a[r].y;  // Re conj(b) a
    v[r] = b[r].x * a[r].y - b[r].y * a[r].x;  // Im conj(b) a
    U += u[r];
    V += v[r];
  }
  for (int i = 0; i < count; ++i) {
    const int type = __ldg(&ops[i].type);
    const int p0 = __ldg(&ops[i].p0), p1 = __ldg(&ops[i].p1);
    const int aux0 = __ldg(&ops[i].aux0), aux1 = __ldg(&ops[i].aux1);
    const float* e = coef + __ldg(&ops[i].coef);
    float val = 0.f;
    if (type == OP_GD_CONST) {
      int sel = (gbase >> aux0) & 1;
      if (aux1 >= 0) sel = 2 * sel + ((gbase >> aux1) & 1);
      const float2 m = ldg2(e + 2 * sel);
      val = m.x * U - m.y * V;
    } else if (type == OP_GD_REG1) {
      const float U1 = sum_bit_dyn<K>(u, p0), V1 = sum_bit_dyn<K>(v, p0);
      const float4 m = ldg4(e);
      val = m.x * (U - U1) - m.y * (V - V1) + m.z * U1 - m.w * V1;
    } else if (type == OP_GD_MIX) {
      const int cb = (gbase >> aux0) & 1;
      const float U1 = sum_bit_dyn<K>(u, p0), V1 = sum_bit_dyn<K>(v, p0);
      const float4 m = ldg4(e + 4 * cb);
      val = m.x * (U - U1) - m.y * (V - V1) + m.z * U1 - m.w * V1;
    } else if (type == OP_GD_REG2) {
      const float Uh = sum_bit_dyn<K>(u, p0), Vh = sum_bit_dyn<K>(v, p0);
      const float Ul = sum_bit_dyn<K>(u, p1), Vl = sum_bit_dyn<K>(v, p1);
      const float U11 = sum_both_dyn<K>(u, p0, p1), V11 = sum_both_dyn<K>(v, p0, p1);
      const float4 m01 = ldg4(e), m23 = ldg4(e + 4);
      val = m01.x * (U - Uh - Ul + U11) - m01.y * (V - Vh - Vl + V11)  // sel 0
            + m01.z * (Ul - U11) - m01.w * (Vl - V11)                   // sel 1: lo bit only
            + m23.x * (Uh - U11) - m23.y * (Vh - V11)                   // sel 2: hi bit only
            + m23.z * U11 - m23.w * V11;                                // sel 3
    }
    scratch[__ldg(&ops[i].gslot) * nthr + tid] = 2.f * val;
  }
}

// ---------------------------------------------------------------------------------
// One pass: smem tile -> registers -> ops -> smem tile.
// BOTH = false: forward (psi only).  BOTH = true: adjoint (psi and lambda, gradients).
// ---------------------------------------------------------------------------------
template <int K, bool BOTH>
__device__ __forceinline__ void run_pass(const KernelArgs& ka, const DevPass* __restrict__ ps, float2* s_psi,
                                         float2* s_lam, uint32_t* s_E, uint32_t goff, uint32_t u) {
  constexpr int R = 1 << K;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  __syncthreads();  // tile complete in smem; s_E reusable
  if (tid < R) {
    uint32_t dep = 0;
#pragma unroll
    for (int j = 0; j < K; ++j)
      if ((tid >> j) & 1) dep |= 1u << __ldg(&ps->regbit[j]);
    s_E[tid] = swz(dep);
  }
  uint32_t base = tid;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int sp = __ldg(&ps->sorted[j]);
    base = ((base >> sp) << (sp + 1)) | (base & ((1u << sp) - 1u));
  }
  const uint32_t B = swz(base);
  const uint32_t gbase = goff | scatter_bits(base, ka.L.runs, ka.L.n_runs);
  __syncthreads();

  float2 a[R];
  float2 b[BOTH ? R : 1];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    a[r] = s_psi[B ^ s_E[r]];
    if constexpr (BOTH) b[r] = s_lam[B ^ s_E[r]];
  }
  const int ngrad = BOTH ? __ldg(&ps->ngrad) : 0;
  float* scratch = reinterpret_cast<float*>(s_psi);
  if (ngrad > 0) __syncthreads();  // every amplitude is in registers: the tiles become scratch

  float2 F = make_float2(1.f, 0.f);
  const int op_end = __ldg(&ps->op_end);
  for (int oi = __ldg(&ps->op_begin); oi < op_end; ++oi) {
    const DevOp* op = ka.ops + oi;
    const int type = __ldg(&op->type);
    const int p0 = __ldg(&op->p0);
    const float* cf = ka.coef + __ldg(&op->coef);
    switch (type) {
      case OP_MAT1: {
        const float4 m0 = ldg4(cf), m1 = ldg4(cf + 4);
        mat1_dyn<K>(a, p0, m0, m1);
        if constexpr (BOTH) mat1_dyn<K>(b, p0, m0, m1);
      } break;
      case OP_MAT2: {
        if (p0 == 0) {
          mat2<K, 0, false>(a, a, cf);
          if constexpr (BOTH) mat2<K, 0, false>(b, b, cf);
        } else {
          mat2<K, 2, false>(a, a, cf);
          if constexpr (BOTH) mat2<K, 2, false>(b, b, cf);
        }
      } break;
      case OP_DCONST_TAB: {
        const uint32_t idx = (gbase >> __ldg(&op->aux0)) & (uint32_t)__ldg(&op->aux1);
        F = cmul(F, ldg2(cf + 2 * idx));
      } break;
      case OP_DCONST_PAIR: {
        const int a0 = __ldg(&op->aux0), a1 = __ldg(&op->aux1);
        int sel = (gbase >> a0) & 1;
        if (a1 >= 0) sel = 2 * sel + ((gbase >> a1) & 1);
        F = cmul(F, ldg2(cf + 2 * sel));
      } break;
      case OP_DREG_TAB: {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float2 c = cmul(F, ldg2(cf + 2 * r));
          a[r] = cmul(a[r], c);
          if constexpr (BOTH) b[r] = cmul(b[r], c);
        }
        F = make_float2(1.f, 0.f);
      } break;
      case OP_DAPPLY: {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          a[r] = cmul(a[r], F);
          if constexpr (BOTH) b[r] = cmul(b[r], F);
        }
        F = make_float2(1.f, 0.f);
      } break;
      case OP_DCROSS: {
        const int cb = (gbase >> __ldg(&op->aux0)) & 1;
        const float4 e = ldg4(cf + 4 * cb);
        const float2 e0 = make_float2(e.x, e.y), e1 = make_float2(e.z, e.w);
        mul_sel_dyn<K>(a, p0, e0, e1);
        if constexpr (BOTH) mul_sel_dyn<K>(b, p0, e0, e1);
      } break;
      default:
        if constexpr (BOTH) {
          if (type == OP_GRAD_MAT1) {
            const float4 m0 = ldg4(cf), m1 = ldg4(cf + 4);
            scratch[__ldg(&op->gslot) * nthr + tid] = grad_mat1_dyn<K>(a, b, p0, m0, m1);
          } else if (type == OP_GRAD_MAT2) {
            const float g = p0 == 0 ? mat2<K, 0, true>(a, b, cf) : mat2<K, 2, true>(a, b, cf);
            scratch[__ldg(&op->gslot) * nthr + tid] = g;
          } else if (type == OP_GD_BEGIN) {
            const int cnt = __ldg(&op->aux0);
            grad_diag_run<K>(a, b, op + 1, cnt, ka.coef, scratch, gbase, tid, nthr);
            oi += cnt;
          }
        }
        break;
    }
  }

  if (ngrad > 0) {
    __syncthreads();
    const uint32_t w = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    const int gs0 = __ldg(&ps->gsym_off);
    double* grow = ka.gacc + (size_t)(ka.per_state ? (ka.grow0 + (int)u) : 0) * ka.P;
    for (int g = w; g < ngrad; g += nw) {
      float s = 0.f;
      for (uint32_t i = lane; i < nthr; i += 32) s += scratch[g * nthr + i];
      s = warp_sum(s);
      if (lane == 0) atomicAdd(grow + __ldg(&ka.gsym[gs0 + g]), (double)s);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    s_psi[B ^ s_E[r]] = a[r];
    if constexpr (BOTH) s_lam[B ^ s_E[r]] = b[r];
  }
}

// ---------------------------------------------------------------------------------
// Expectation phase: E_j = Re <psi|H_j|psi>, and (adjoint) lambda = sum_j g_j H_j psi.
// H psi[i] = sum_groups coefficient_g(i) psi[i ^ x_g]; all terms of a group share x.
// ---------------------------------------------------------------------------------
template <int K, bool ADJ>
__device__ __forceinline__ void expect_phase(const KernelArgs& ka, float2* s_psi, float2* s_lam, uint32_t goff,
                                             uint32_t u, const float2* __restrict__ psi_u) {
  constexpr int R = 1 << K;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  __syncthreads();
  float2 a[R];
  float2 lam[ADJ ? R : 1];
  const uint32_t gi_tid = goff | scatter_bits(tid, ka.L.runs, ka.L.n_runs);
#pragma unroll
  for (int m = 0; m < R; ++m) {
    a[m] = s_psi[swz((uint32_t)m * nthr | tid)];
    if constexpr (ADJ) lam[m] = make_float2(0.f, 0.f);
  }
  const bool want_lam = ADJ && ka.dgrad != nullptr;
  for (int j = 0; j < ka.O; ++j) {
    const float gj = want_lam ? __ldg(&ka.dgrad[(size_t)u * ka.O + j]) : 0.f;
    float ej = 0.f;
    const int g_end = __ldg(&ka.opranges[j].group_end);
    for (int g = __ldg(&ka.opranges[j].group_begin); g < g_end; ++g) {
      const uint32_t x = __ldg(&ka.groups[g].x);
      const int xl = __ldg(&ka.groups[g].xl);
      const int t0 = __ldg(&ka.groups[g].term_begin), t1 = __ldg(&ka.groups[g].term_end);
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const uint32_t l = (uint32_t)m * nthr | tid;
        const uint32_t gi = gi_tid | scatter_bits((uint32_t)m * nthr, ka.L.runs, ka.L.n_runs);
        float cr = 0.f, ci = 0.f;
        for (int t = t0; t < t1; ++t) {
          const float4 tv = __ldg(reinterpret_cast<const float4*>(ka.terms + t));
          const uint32_t sgn = (uint32_t)(__popc(gi & __float_as_uint(tv.z)) & 1) << 31;
          cr += __uint_as_float(__float_as_uint(tv.x) ^ sgn);
          ci += __uint_as_float(__float_as_uint(tv.y) ^ sgn);
        }
        const float2 p = xl >= 0 ? s_psi[swz(l ^ (uint32_t)xl)] : psi_u[gi ^ x];
        const float hr = cr * p.x - ci * p.y, hi = cr * p.y + ci * p.x;
        ej += a[m].x * hr + a[m].y * hi;
        if constexpr (ADJ) {
          lam[m].x += gj * hr;
          lam[m].y += gj * hi;
        }
      }
    }
    ej = warp_sum(ej);
    if ((tid & 31) == 0) atomicAdd(&ka.eacc[(size_t)u * ka.O + j], (double)ej);
  }
  if constexpr (ADJ) {
#pragma unroll
    for (int m = 0; m < R; ++m) s_lam[swz((uint32_t)m * nthr | tid)] = lam[m];
  }
}

template <int K>
__device__ __forceinline__ void load_tile(float2* s, const float2* __restrict__ g, uint32_t goff, const KernelArgs& ka) {
  constexpr int R = 1 << K;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  const uint32_t gt = goff | scatter_bits(tid, ka.L.runs, ka.L.n_runs);
  float2 v[R];
#pragma unroll
  for (int m = 0; m < R; ++m) v[m] = g[gt | scatter_bits((uint32_t)m * nthr, ka.L.runs, ka.L.n_runs)];
#pragma unroll
  for (int m = 0; m < R; ++m) s[swz((uint32_t)m * nthr | tid)] = v[m];
}
template <int K>
__device__ __forceinline__ void store_tile(const float2* s, float2* __restrict__ g, uint32_t goff, const KernelArgs& ka) {
  constexpr int R = 1 << K;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  const uint32_t gt = goff | scatter_bits(tid, ka.L.runs, ka.L.n_runs);
#pragma unroll
  for (int m = 0; m < R; ++m)
    g[gt | scatter_bits((uint32_t)m * nthr, ka.L.runs, ka.L.n_runs)] = s[swz((uint32_t)m * nthr | tid)];
}

template <int K, bool ADJ>
constexpr int sweep_max_threads() { return ADJ ? (1 << (13 - K)) : 512; }

// grid = chunk * 2^(n-T) CTAs, block = 2^(T-K) threads, dynamic smem = (ADJ ? 2 : 1) * 8 * 2^T bytes.
template <int K, bool ADJ>
__global__ void __launch_bounds__(sweep_max_threads<K, ADJ>()) sweep_kernel(const __grid_constant__ KernelArgs ka) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_E[1 << kMaxRegQubits];
  float2* s_psi = reinterpret_cast<float2*>(smem_raw);
  float2* s_lam = s_psi + (ADJ ? (1u << ka.T) : 0u);
  constexpr int R = 1 << K;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  const uint32_t tshift = (uint32_t)(ka.n - ka.T);
  const uint32_t u = blockIdx.x >> tshift;
  const uint32_t tile = blockIdx.x & ((1u << tshift) - 1u);
  const uint32_t goff = scatter_bits(tile, ka.L.oruns, ka.L.n_oruns);
  float2* psi_u = ka.psi ? ka.psi + ((size_t)u << ka.n) : nullptr;
  float2* lam_u = ka.lam ? ka.lam + ((size_t)u << ka.n) : nullptr;
  const uint32_t flags = ka.L.flags;

  bool active = true;
  if (flags & LF_INIT_BASIS) {
    const uint32_t basis = (uint32_t)ka.basis[u];
    active = (basis & ~ka.L.tile_mask) == goff;
    const uint32_t lb = gather_bits(basis, ka.L.runs, ka.L.n_runs);
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const uint32_t l = (uint32_t)m * nthr | tid;
      s_psi[swz(l)] = make_float2((active && l == lb) ? 1.f : 0.f, 0.f);
    }
  } else if (flags & LF_LOAD_PSI) {
    load_tile<K>(s_psi, psi_u, goff, ka);
  }
  if constexpr (ADJ) {
    if (flags & LF_LOAD_LAM) load_tile<K>(s_lam, lam_u, goff, ka);
  }

  if (active) {
    for (int p = ka.L.pass_a_begin; p < ka.L.pass_a_end; ++p)
      run_pass<K, false>(ka, ka.passes + p, s_psi, s_lam, s_E, goff, u);
  }
  if (flags & LF_WRITE_STATE) {
    __syncthreads();
    store_tile<K>(s_psi, ka.state_out + ((size_t)u << ka.n), goff, ka);
  }
  if (flags & LF_EXPECT) expect_phase<K, ADJ>(ka, s_psi, s_lam, goff, u, psi_u);
  if constexpr (ADJ) {
    for (int p = ka.L.pass_b_begin; p < ka.L.pass_b_end; ++p)
      run_pass<K, true>(ka, ka.passes + p, s_psi, s_lam, s_E, goff, u);
  }
  if (flags & (LF_STORE_PSI | LF_STORE_LAM)) {
    __syncthreads();
    if (flags & LF_STORE_PSI) store_tile<K>(s_psi, psi_u, goff, ka);
    if constexpr (ADJ) {
      if (flags & LF_STORE_LAM) store_tile<K>(s_lam, lam_u, goff, ka);
    }
  }
}

// ---------------------------------------------------------------------------------
// Coefficient preparation: symbols -> gate matrices, gradient matrices, phase tables.
// One CTA per job; float64 math, float32 results.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void write_c(float* out, int i, cd v) {
  out[2 * i] = (float)v.re;
  out[2 * i + 1] = (float)v.im;
}

constexpr int kPrepThreads = 128;
constexpr int kPrepBatch = 64;

__global__ void __launch_bounds__(kPrepThreads) prep_kernel(const PrepJob* __restrict__ jobs,
                                                            const int32_t* __restrict__ lists,
                                                            const qhbm_gate_t* __restrict__ gates,
                                                            const float* __restrict__ symbols,
                                                            float* __restrict__ coef, int mode) {
  const PrepJob job = jobs[blockIdx.x];
  const int32_t* list = lists + job.list_off;
  float* out = coef + job.out;
  const int tid = threadIdx.x;
  if (job.kind == PJ_DTAB) {
    __shared__ cd s_diag[kPrepBatch][4];
    __shared__ int s_pos[kPrepBatch][2];
    const int entries = 1 << job.d;
    const int ntrip = job.list_len / 3;
    cd acc[2];  // up to 2 entries per thread (256 entries / 128 threads)
    acc[0] = mk(1, 0);
    acc[1] = mk(1, 0);
    for (int b0 = 0; b0 < ntrip; b0 += kPrepBatch) {
      const int nb = min(kPrepBatch, ntrip - b0);
      __syncthreads();
      if (tid < nb) {
        const int gi = list[3 * (b0 + tid)];
        cd m[16];
        const int dim = gate_matrix_of(gates[gi], symbols, m);
        for (int i = 0; i < 4; ++i) {
          cd v = i < dim ? m[i * dim + i] : mk(1, 0);
          s_diag[tid][i] = job.a ? conj(v) : v;
        }
        s_pos[tid][0] = list[3 * (b0 + tid) + 1];
        s_pos[tid][1] = list[3 * (b0 + tid) + 2];
      }
      __syncthreads();
      for (int e = 0; e < 2; ++e) {
        const int v = tid + e * kPrepThreads;
        if (v >= entries) break;
        for (int t = 0; t < nb; ++t) {
          int sel = (v >> s_pos[t][0]) & 1;
          if (s_pos[t][1] >= 0) sel = 2 * sel + ((v >> s_pos[t][1]) & 1);
          acc[e] = acc[e] * s_diag[t][sel];
        }
      }
    }
    for (int e = 0; e < 2; ++e) {
      const int v = tid + e * kPrepThreads;
      if (v < entries) write_c(out, v, acc[e]);
    }
    return;
  }
  if (tid != 0) return;
  cd m[16], t[16], w[16];
  switch (job.kind) {
    case PJ_MAT1: {
      cd acc[4] = {mk(1, 0), mk(0, 0), mk(0, 0), mk(1, 0)};
      for (int i = 0; i < job.list_len; ++i) {
        gate_matrix_of(gates[list[i]], symbols, m);
        matmul(m, acc, 2, t);
        for (int k = 0; k < 4; ++k) acc[k] = t[k];
      }
      if (job.a) { dagger(acc, 2, t); for (int k = 0; k < 4; ++k) acc[k] = t[k]; }
      for (int k = 0; k < 4; ++k) write_c(out, k, acc[k]);
    } break;
    case PJ_MAT2: {
      gate_matrix_of(gates[list[0]], symbols, m);
      if (job.b) { swap_qubits(m, t); for (int k = 0; k < 16; ++k) m[k] = t[k]; }
      if (job.a) { dagger(m, 4, t); for (int k = 0; k < 16; ++k) m[k] = t[k]; }
      for (int k = 0; k < 16; ++k) write_c(out, k, m[k]);
    } break;
    case PJ_GRAD1:
    case PJ_GRAD2:
    case PJ_GDIAG: {
      const qhbm_gate_t g = gates[list[0]];
      const int dim = gate_matrix_of(g, symbols, m);
      gate_derivative(g, symbols, job.c, mode, t);
      dagger(m, dim, w);
      matmul(t, w, dim, m);  // M = dG G^dagger
      if (dim == 4 && job.b) { swap_qubits(m, t); for (int k = 0; k < 16; ++k) m[k] = t[k]; }
      if (job.kind == PJ_GDIAG) {
        for (int k = 0; k < 4; ++k) write_c(out, k, k < dim ? m[k * dim + k] : mk(0, 0));
      } else {
        for (int k = 0; k < dim * dim; ++k) write_c(out, k, m[k]);
      }
    } break;
    case PJ_DPAIR: {
      const int dim = gate_matrix_of(gates[list[0]], symbols, m);
      if (dim == 4 && job.b) { swap_qubits(m, t); for (int k = 0; k < 16; ++k) m[k] = t[k]; }
      for (int k = 0; k < 4; ++k) {
        cd v = k < dim ? m[k * dim + k] : mk(1, 0);
        write_c(out, k, job.a ? conj(v) : v);
      }
    } break;
    default: break;
  }
}

__global__ void finalize_kernel(const double* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}

}  // namespace qhbm
