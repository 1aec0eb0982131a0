/* tfq_cpu.c -- CPU restatement (complex64, OpenMP) of the algorithm TFQ 0.6.1 runs behind
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
 * Gate matrices (G, G^dag blocks, dG) are computed by oracle/qhbm_oracle.py and passed in. */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef float _Complex c64;

static void apply1(c64* s, int n, int q, const c64* m) {
  const size_t stride = (size_t)1 << (n - 1 - q);
  const size_t N = (size_t)1 << n;
  for (size_t base = 0; base < N; base += 2 * stride)
    for (size_t k = 0; k < stride; ++k) {
      const c64 a0 = s[base + k], a1 = s[base + k + stride];
      s[base + k] = m[0] * a0 + m[1] * a1;
      s[base + k + stride] = m[2] * a0 + m[3] * a1;
    }
}

/* m is big-endian in (q0, q1): index = 2*bit(q0) + bit(q1) */
static void apply2(c64* s, int n, int q0, int q1, const c64* m) {
  const size_t s0 = (size_t)1 << (n - 1 - q0), s1 = (size_t)1 << (n - 1 - q1);
  const size_t N = (size_t)1 << n;
  const size_t hi = s0 > s1 ? s0 : s1, lo = s0 > s1 ? s1 : s0;
  for (size_t a = 0; a < N; a += 2 * hi)
    for (size_t b = 0; b < hi; b += 2 * lo)
      for (size_t k = 0; k < lo; ++k) {
        const size_t i = a + b + k;
        const c64 x0 = s[i], x1 = s[i + s1], x2 = s[i + s0], x3 = s[i + s0 + s1];
        s[i] = m[0] * x0 + m[1] * x1 + m[2] * x2 + m[3] * x3;
        s[i + s1] = m[4] * x0 + m[5] * x1 + m[6] * x2 + m[7] * x3;
        s[i + s0] = m[8] * x0 + m[9] * x1 + m[10] * x2 + m[11] * x3;
        s[i + s0 + s1] = m[12] * x0 + m[13] * x1 + m[14] * x2 + m[15] * x3;
      }
}

static void apply_block(c64* s, int n, int q0, int q1, const c64* m) {
  if (q1 < 0) apply1(s, n, q0, m);
  else apply2(s, n, q0, q1, m);
}

/* out += g * coeff * P psi  (P given by x/z masks over index bits; Y = x&z) and returns
 * coeff * Re<psi|P|psi>. */
static double pauli_term(const c64* psi, c64* out, int n, float coeff, uint32_t x, uint32_t z, float g) {
  const size_t N = (size_t)1 << n;
  const int ny = __builtin_popcount(x & z) & 3;
  c64 k = coeff;
  if (ny == 1) k = -I * coeff;
  else if (ny == 2) k = -coeff;
  else if (ny == 3) k = I * coeff;
  double acc = 0.0;
  for (size_t i = 0; i < N; ++i) {
    const c64 h = ((__builtin_popcount((uint32_t)i & z) & 1) ? -k : k) * psi[i ^ x];
    acc += (double)(crealf(psi[i]) * crealf(h) + cimagf(psi[i]) * cimagf(h));
    if (out) out[i] += g * h;
  }
  return acc;
}

typedef struct {
  int n, n_blocks, n_gates, n_ops, n_sym;
  const int32_t* bq0; const int32_t* bq1; const c64* bmat;      /* fused forward blocks, 16 entries each */
  const int32_t* gq0; const int32_t* gq1; const c64* gdag;      /* per gate: G^dagger, 16 entries each   */
  const int32_t* grad_off;                                      /* n_gates+1 offsets into grad lists      */
  const int32_t* grad_sym; const c64* grad_mat;                 /* dG per (gate, symbol), 16 entries      */
  const float* t_coeff; const uint32_t* t_x; const uint32_t* t_z; const int32_t* t_off; /* n_ops+1 */
} problem_t;

static void forward(const problem_t* p, c64* psi, uint64_t basis) {
  memset(psi, 0, sizeof(c64) << p->n);
  psi[basis] = 1.0f;
  for (int b = 0; b < p->n_blocks; ++b) apply_block(psi, p->n, p->bq0[b], p->bq1[b], p->bmat + 16 * (size_t)b);
}

/* f32[U,O] expectations (TfqSimulateExpectation). */
int tfq_cpu_expectation(const problem_t* p, const uint64_t* basis, int64_t U, float* out, int threads) {
  int fail = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
  {
    c64* psi = (c64*)malloc(sizeof(c64) << p->n);
    if (!psi) fail = 1;
#pragma omp for schedule(dynamic, 1)
    for (int64_t u = 0; u < U; ++u) {
      if (!psi) continue;
      forward(p, psi, basis[u]);
      for (int j = 0; j < p->n_ops; ++j) {
        double e = 0.0;
        for (int t = p->t_off[j]; t < p->t_off[j + 1]; ++t)
          e += pauli_term(psi, NULL, p->n, p->t_coeff[t], p->t_x[t], p->t_z[t], 0.f);
        out[u * p->n_ops + j] = (float)e;
      }
    }
    free(psi);
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
    c64* psi = (c64*)malloc(sizeof(c64) * N);
    c64* lam = (c64*)malloc(sizeof(c64) * N);
    c64* scr = (c64*)malloc(sizeof(c64) * N);
    if (!psi || !lam || !scr) fail = 1;
#pragma omp for schedule(dynamic, 1)
    for (int64_t u = 0; u < U; ++u) {
      if (fail) continue;
      forward(p, psi, basis[u]);
      memset(lam, 0, sizeof(c64) * N);
      for (int j = 0; j < p->n_ops; ++j) {
        const float g = dgrad[u * p->n_ops + j];
        double e = 0.0;
        for (int t = p->t_off[j]; t < p->t_off[j + 1]; ++t)
          e += pauli_term(psi, g != 0.f ? lam : NULL, p->n, p->t_coeff[t], p->t_x[t], p->t_z[t], g);
        out[u * p->n_ops + j] = (float)e;
      }
      float* gr = grad + u * p->n_sym;
      for (int s = 0; s < p->n_sym; ++s) gr[s] = 0.f;
      for (int k = p->n_gates - 1; k >= 0; --k) {
        apply_block(psi, p->n, p->gq0[k], p->gq1[k], p->gdag + 16 * (size_t)k);
        for (int gi = p->grad_off[k]; gi < p->grad_off[k + 1]; ++gi) {
          memcpy(scr, psi, sizeof(c64) * N);
          apply_block(scr, p->n, p->gq0[k], p->gq1[k], p->grad_mat + 16 * (size_t)gi);
          double acc = 0.0;
          for (size_t i = 0; i < N; ++i)
            acc += (double)(crealf(lam[i]) * crealf(scr[i]) + cimagf(lam[i]) * cimagf(scr[i]));
          gr[p->grad_sym[gi]] += (float)(2.0 * acc);
        }
        apply_block(lam, p->n, p->gq0[k], p->gq1[k], p->gdag + 16 * (size_t)k);
      }
    }
    free(psi); free(lam); free(scr);
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
