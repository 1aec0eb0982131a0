/* tfq_cpu.c -- CPU restatement (complex64, OpenMP + SIMD) of the algorithm TFQ 0.6.1 runs behind
 * tfq.layers.Expectation() + its adjoint differentiator, i.e. the C++ ops
 * TfqSimulateExpectation and TfqAdjointGradient that qhbmlib reaches from
 * /root/reference/qhbmlib/inference/qnn.py:112,134-138.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY ("CPU restatement of TFQ 0.6.1's algorithm -- not TFQ
 * itself"; TFQ is not installable here).  Used by tests/ as a second checker and by bench.py
 * as the cpu_baseline / --impl reference arm.  Never linked into libqhbm_b200.so.
 *
 * Algorithm restated (SURVEY.md sections 2.2, A.5, A.6):
 *   forward : state-vector simulation from |basis>, gates pre-fused into <=2-qubit blocks
 *             (qsim BasicGateFuser), one state per thread, parallel-for over circuits;
 *   <H>     : term by term, sum_t coeff_t Re<psi|P_t|psi>, float32 state, double accumulation;
 *   adjoint : re-simulate forward, lambda = sum_j g_j H_j psi, walk gates in reverse:
 *             psi <- G^dag psi; grad[s] += 2 Re<lambda| dG |psi>; lambda <- G^dag lambda.
 * Gate matrices (G, G^dag blocks, dG) are computed by oracle/qhbm_oracle.py and passed in.
 *
 * Round 2: the state is stored as split real / imaginary float arrays and every hot loop is written
 * in real arithmetic over a contiguous inner index with `omp simd`, so that gcc vectorises it (AVX /
 * AVX2 / AVX-512 as -march=native offers), the way qsim's hand-written SIMD kernels do.  Gates whose
 * inner stride is below one SIMD vector (the three lowest index bits) run a scalar loop -- qsim
 * shuffles inside registers there; that remaining gap is stated next to every CPU number.  The
 * gradient inner product <lambda| dG |psi> is evaluated on the fly (no scratch copy of the state). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float re, im; } c64;  /* layout of numpy complex64 */

#define SIMD_MIN 8

/* 2x2 block on qubit q: (a0, a1) <- m (a0, a1) for every pair of amplitudes differing in bit (n-1-q). */
static void apply1(float* restrict re, float* restrict im, int n, int q, const c64* m) {
  const size_t stride = (size_t)1 << (n - 1 - q);
  const size_t N = (size_t)1 << n;
  const float m00r = m[0].re, m00i = m[0].im, m01r = m[1].re, m01i = m[1].im;
  const float m10r = m[2].re, m10i = m[2].im, m11r = m[3].re, m11i = m[3].im;
  for (size_t base = 0; base < N; base += 2 * stride) {
    float* restrict r0 = re + base; float* restrict i0 = im + base;
    float* restrict r1 = re + base + stride; float* restrict i1 = im + base + stride;
#pragma omp simd
    for (size_t k = 0; k < stride; ++k) {
      const float ar = r0[k], ai = i0[k], br = r1[k], bi = i1[k];
      r0[k] = m00r * ar - m00i * ai + m01r * br - m01i * bi;
      i0[k] = m00r * ai + m00i * ar + m01r * bi + m01i * br;
      r1[k] = m10r * ar - m10i * ai + m11r * br - m11i * bi;
      i1[k] = m10r * ai + m10i * ar + m11r * bi + m11i * br;
    }
  }
}

/* 4x4 block, m big-endian in (q0, q1): row/column index = 2*bit(q0) + bit(q1). */
static void apply2(float* restrict re, float* restrict im, int n, int q0, int q1, const c64* m) {
  const size_t s0 = (size_t)1 << (n - 1 - q0), s1 = (size_t)1 << (n - 1 - q1);
  const size_t N = (size_t)1 << n;
  const size_t hi = s0 > s1 ? s0 : s1, lo = s0 > s1 ? s1 : s0;
  float mr[16], mi[16];
  for (int k = 0; k < 16; ++k) { mr[k] = m[k].re; mi[k] = m[k].im; }
  for (size_t a = 0; a < N; a += 2 * hi)
    for (size_t b = 0; b < hi; b += 2 * lo) {
      const size_t o = a + b;
      float* restrict xr[4] = {re + o, re + o + s1, re + o + s0, re + o + s0 + s1};
      float* restrict xi[4] = {im + o, im + o + s1, im + o + s0, im + o + s0 + s1};
#pragma omp simd
      for (size_t k = 0; k < lo; ++k) {
        const float ar[4] = {xr[0][k], xr[1][k], xr[2][k], xr[3][k]};
        const float ai[4] = {xi[0][k], xi[1][k], xi[2][k], xi[3][k]};
        for (int r = 0; r < 4; ++r) {
          float yr = 0.f, yi = 0.f;
          for (int c = 0; c < 4; ++c) {
            yr += mr[4 * r + c] * ar[c] - mi[4 * r + c] * ai[c];
            yi += mr[4 * r + c] * ai[c] + mi[4 * r + c] * ar[c];
          }
          xr[r][k] = yr;
          xi[r][k] = yi;
        }
      }
    }
}

static void apply_block(float* re, float* im, int n, int q0, int q1, const c64* m) {
  if (q1 < 0) apply1(re, im, n, q0, m);
  else apply2(re, im, n, q0, q1, m);
}

/* 2 Re <lam| D |psi> for a 2x2 / 4x4 block D, without materialising D psi. */
static double grad1(const float* restrict pr, const float* restrict pi, const float* restrict lr,
                    const float* restrict li, int n, int q, const c64* d) {
  const size_t stride = (size_t)1 << (n - 1 - q);
  const size_t N = (size_t)1 << n;
  const float d00r = d[0].re, d00i = d[0].im, d01r = d[1].re, d01i = d[1].im;
  const float d10r = d[2].re, d10i = d[2].im, d11r = d[3].re, d11i = d[3].im;
  double total = 0.0;
  for (size_t base = 0; base < N; base += 2 * stride) {
    float acc = 0.f;
#pragma omp simd reduction(+ : acc)
    for (size_t k = 0; k < stride; ++k) {
      const size_t i0 = base + k, i1 = base + k + stride;
      const float ar = pr[i0], ai = pi[i0], br = pr[i1], bi = pi[i1];
      const float y0r = d00r * ar - d00i * ai + d01r * br - d01i * bi;
      const float y0i = d00r * ai + d00i * ar + d01r * bi + d01i * br;
      const float y1r = d10r * ar - d10i * ai + d11r * br - d11i * bi;
      const float y1i = d10r * ai + d10i * ar + d11r * bi + d11i * br;
      acc += lr[i0] * y0r + li[i0] * y0i + lr[i1] * y1r + li[i1] * y1i;
    }
    total += (double)acc;
  }
  return 2.0 * total;
}

static double grad2(const float* restrict pr, const float* restrict pi, const float* restrict lr,
                    const float* restrict li, int n, int q0, int q1, const c64* d) {
  const size_t s0 = (size_t)1 << (n - 1 - q0), s1 = (size_t)1 << (n - 1 - q1);
  const size_t N = (size_t)1 << n;
  const size_t hi = s0 > s1 ? s0 : s1, lo = s0 > s1 ? s1 : s0;
  float mr[16], mi[16];
  for (int k = 0; k < 16; ++k) { mr[k] = d[k].re; mi[k] = d[k].im; }
  const size_t off[4] = {0, s1, s0, s0 + s1};
  double total = 0.0;
  for (size_t a = 0; a < N; a += 2 * hi)
    for (size_t b = 0; b < hi; b += 2 * lo) {
      const size_t o = a + b;
      float acc = 0.f;
#pragma omp simd reduction(+ : acc)
      for (size_t k = 0; k < lo; ++k) {
        float ar[4], ai[4];
        for (int c = 0; c < 4; ++c) { ar[c] = pr[o + off[c] + k]; ai[c] = pi[o + off[c] + k]; }
        for (int r = 0; r < 4; ++r) {
          float yr = 0.f, yi = 0.f;
          for (int c = 0; c < 4; ++c) {
            yr += mr[4 * r + c] * ar[c] - mi[4 * r + c] * ai[c];
            yi += mr[4 * r + c] * ai[c] + mi[4 * r + c] * ar[c];
          }
          acc += lr[o + off[r] + k] * yr + li[o + off[r] + k] * yi;
        }
      }
      total += (double)acc;
    }
  return 2.0 * total;
}

/* out += g * coeff * P psi  (P given by x/z masks over index bits; Y = x&z) and returns
 * coeff * Re<psi|P|psi>.  Blocks of SIMD_MIN consecutive indices share the sign of the high index
 * bits; inside a block the sign pattern and the partner permutation come from small tables. */
static double pauli_term(const float* restrict pr, const float* restrict pi, float* restrict outr,
                         float* restrict outi, int n, float coeff, uint32_t x, uint32_t z, float g) {
  const size_t N = (size_t)1 << n;
  const int ny = __builtin_popcount(x & z) & 3;
  /* k = coeff * (-i)^ny */
  float kr = 0.f, ki = 0.f;
  if (ny == 0) kr = coeff; else if (ny == 1) ki = -coeff; else if (ny == 2) kr = -coeff; else ki = coeff;
  const size_t B = N < SIMD_MIN ? N : SIMD_MIN;
  float lane_sign[SIMD_MIN];
  size_t perm[SIMD_MIN];
  for (size_t j = 0; j < B; ++j) {
    lane_sign[j] = (__builtin_popcount((uint32_t)j & z) & 1) ? -1.f : 1.f;
    perm[j] = j ^ (x & (B - 1));
  }
  const uint32_t xhi = x & ~(uint32_t)(B - 1);
  const int contiguous = (x & (B - 1)) == 0;
  double total = 0.0;
  for (size_t i0 = 0; i0 < N; i0 += B) {
    const float hs = (__builtin_popcount((uint32_t)i0 & z) & 1) ? -1.f : 1.f;
    const size_t p0 = i0 ^ xhi;
    float acc = 0.f;
    if (contiguous) {
#pragma omp simd reduction(+ : acc)
      for (size_t j = 0; j < B; ++j) {
        const float s = hs * lane_sign[j];
        const float hr = s * (kr * pr[p0 + j] - ki * pi[p0 + j]);
        const float hi = s * (kr * pi[p0 + j] + ki * pr[p0 + j]);
        acc += pr[i0 + j] * hr + pi[i0 + j] * hi;
        if (outr) { outr[i0 + j] += g * hr; outi[i0 + j] += g * hi; }
      }
    } else {
      for (size_t j = 0; j < B; ++j) {
        const float s = hs * lane_sign[j];
        const size_t p = p0 + perm[j];
        const float hr = s * (kr * pr[p] - ki * pi[p]);
        const float hi = s * (kr * pi[p] + ki * pr[p]);
        acc += pr[i0 + j] * hr + pi[i0 + j] * hi;
        if (outr) { outr[i0 + j] += g * hr; outi[i0 + j] += g * hi; }
      }
    }
    total += (double)acc;
  }
  return total;
}

typedef struct {
  int n, n_blocks, n_gates, n_ops, n_sym;
  const int32_t* bq0; const int32_t* bq1; const c64* bmat;      /* fused forward blocks, 16 entries each */
  const int32_t* gq0; const int32_t* gq1; const c64* gdag;      /* per gate: G^dagger, 16 entries each   */
  const int32_t* grad_off;                                      /* n_gates+1 offsets into grad lists      */
  const int32_t* grad_sym; const c64* grad_mat;                 /* dG per (gate, symbol), 16 entries      */
  const float* t_coeff; const uint32_t* t_x; const uint32_t* t_z; const int32_t* t_off; /* n_ops+1 */
} problem_t;

static void forward(const problem_t* p, float* re, float* im, uint64_t basis) {
  memset(re, 0, sizeof(float) << p->n);
  memset(im, 0, sizeof(float) << p->n);
  re[basis] = 1.0f;
  for (int b = 0; b < p->n_blocks; ++b) apply_block(re, im, p->n, p->bq0[b], p->bq1[b], p->bmat + 16 * (size_t)b);
}

/* f32[U,O] expectations (TfqSimulateExpectation). */
int tfq_cpu_expectation(const problem_t* p, const uint64_t* basis, int64_t U, float* out, int threads) {
  int fail = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
  {
    const size_t N = (size_t)1 << p->n;
    float* buf = (float*)aligned_alloc(64, sizeof(float) * 2 * N);
    if (!buf) fail = 1;
#pragma omp for schedule(dynamic, 1)
    for (int64_t u = 0; u < U; ++u) {
      if (!buf) continue;
      float* re = buf; float* im = buf + N;
      forward(p, re, im, basis[u]);
      for (int j = 0; j < p->n_ops; ++j) {
        double e = 0.0;
        for (int t = p->t_off[j]; t < p->t_off[j + 1]; ++t)
          e += pauli_term(re, im, NULL, NULL, p->n, p->t_coeff[t], p->t_x[t], p->t_z[t], 0.f);
        out[u * p->n_ops + j] = (float)e;
      }
    }
    free(buf);
  }
  return fail;
}

/* f32[U,O] expectations and f32[U,P] gradients (TfqAdjointGradient; forward re-simulated). */
int tfq_cpu_adjoint(const problem_t* p, const uint64_t* basis, int64_t U, const float* dgrad, float* out,
                    float* grad, int threads) {
  int fail = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
  {
    const size_t N = (size_t)1 << p->n;
    float* buf = (float*)aligned_alloc(64, sizeof(float) * 4 * N);
    if (!buf) fail = 1;
#pragma omp for schedule(dynamic, 1)
    for (int64_t u = 0; u < U; ++u) {
      if (!buf) continue;
      float* pr = buf; float* pi = buf + N; float* lr = buf + 2 * N; float* li = buf + 3 * N;
      forward(p, pr, pi, basis[u]);
      memset(lr, 0, sizeof(float) * 2 * N);
      for (int j = 0; j < p->n_ops; ++j) {
        const float g = dgrad[u * p->n_ops + j];
        double e = 0.0;
        for (int t = p->t_off[j]; t < p->t_off[j + 1]; ++t)
          e += pauli_term(pr, pi, g != 0.f ? lr : NULL, li, p->n, p->t_coeff[t], p->t_x[t], p->t_z[t], g);
        out[u * p->n_ops + j] = (float)e;
      }
      float* gr = grad + u * p->n_sym;
      for (int s = 0; s < p->n_sym; ++s) gr[s] = 0.f;
      for (int k = p->n_gates - 1; k >= 0; --k) {
        apply_block(pr, pi, p->n, p->gq0[k], p->gq1[k], p->gdag + 16 * (size_t)k);
        for (int gi = p->grad_off[k]; gi < p->grad_off[k + 1]; ++gi) {
          const c64* d = p->grad_mat + 16 * (size_t)gi;
          const double v = p->gq1[k] < 0 ? grad1(pr, pi, lr, li, p->n, p->gq0[k], d)
                                         : grad2(pr, pi, lr, li, p->n, p->gq0[k], p->gq1[k], d);
          gr[p->grad_sym[gi]] += (float)v;
        }
        apply_block(lr, li, p->n, p->gq0[k], p->gq1[k], p->gdag + 16 * (size_t)k);
      }
    }
    free(buf);
  }
  return fail;
}

int tfq_cpu_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
