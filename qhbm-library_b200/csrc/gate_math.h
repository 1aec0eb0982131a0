// gate_math.h -- gate matrices in double precision, usable from host and device.
//
// Definitions follow cirq 0.14.1 (the gate semantics TFQ 0.6.1 serialises; the
// reference reaches them through qhbmlib/inference/qnn.py:112,134-138).  See
// SURVEY.md App. A.4.  Two-qubit matrices are big-endian in (q0, q1).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/qhbm_b200.h"

#if defined(__CUDACC__)
#define QHBM_HD __host__ __device__ __forceinline__
#else
#define QHBM_HD inline
#endif

namespace qhbm {

struct cd {
  double re, im;
};
QHBM_HD cd mk(double r, double i) { cd c; c.re = r; c.im = i; return c; }
QHBM_HD cd operator+(cd a, cd b) { return mk(a.re + b.re, a.im + b.im); }
QHBM_HD cd operator-(cd a, cd b) { return mk(a.re - b.re, a.im - b.im); }
QHBM_HD cd operator*(cd a, cd b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
QHBM_HD cd operator*(double s, cd a) { return mk(s * a.re, s * a.im); }
QHBM_HD cd conj(cd a) { return mk(a.re, -a.im); }
// e^{i pi t}
QHBM_HD cd expipi(double t) {
  double s, c;
#if defined(__CUDA_ARCH__)
  sincospi(t, &s, &c);
#else
  const double kPi = 3.14159265358979323846;
  s = sin(kPi * t);
  c = cos(kPi * t);
#endif
  return mk(c, s);
}
QHBM_HD cd expi(double x) { return mk(cos(x), sin(x)); }

QHBM_HD bool gate_is_two_qubit(int type) {
  return type == QHBM_GATE_CZPOW || type == QHBM_GATE_CNOTPOW || type == QHBM_GATE_SWAPPOW ||
         type == QHBM_GATE_ISWAPPOW || type == QHBM_GATE_XXPOW || type == QHBM_GATE_YYPOW ||
         type == QHBM_GATE_ZZPOW || type == QHBM_GATE_FSIM || type == QHBM_GATE_PHASEDISWAPPOW;
}
QHBM_HD bool gate_is_diagonal(int type) {
  return type == QHBM_GATE_I || type == QHBM_GATE_ZPOW || type == QHBM_GATE_CZPOW ||
         type == QHBM_GATE_ZZPOW;
}
QHBM_HD int gate_num_params(int type) {
  if (type == QHBM_GATE_I) return 0;
  if (type == QHBM_GATE_PHASEDXPOW || type == QHBM_GATE_FSIM || type == QHBM_GATE_PHASEDISWAPPOW)
    return 2;
  return 1;
}

// m (row-major, dim x dim, dim = 2 or 4) = e^{i pi t g} (I + (e^{i pi t} - 1) P1)
// where P1 is given as a real/complex projector.
QHBM_HD void two_level(const cd* p1, int dim, double t, double g, cd* m) {
  cd ph = expipi(t * g);
  cd e1 = expipi(t) - mk(1.0, 0.0);
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < dim; ++j) {
      cd v = e1 * p1[i * dim + j];
      if (i == j) v = v + mk(1.0, 0.0);
      m[i * dim + j] = ph * v;
    }
}

// Returns the matrix dimension (2 or 4); m must hold 16 entries.
QHBM_HD int gate_matrix(int type, const double* p, double g, cd* m) {
  const double h = 0.5;
  const double r = 0.70710678118654752440;
  cd z = mk(0, 0), one = mk(1, 0);
  cd P[16];
  for (int i = 0; i < 16; ++i) { P[i] = z; m[i] = z; }
  switch (type) {
    case QHBM_GATE_I:
      m[0] = one; m[3] = one;
      return 2;
    case QHBM_GATE_XPOW:  // P1 = (I - X)/2
      P[0] = mk(h, 0); P[1] = mk(-h, 0); P[2] = mk(-h, 0); P[3] = mk(h, 0);
      two_level(P, 2, p[0], g, m);
      return 2;
    case QHBM_GATE_YPOW:  // (I - Y)/2, Y = [[0,-i],[i,0]]
      P[0] = mk(h, 0); P[1] = mk(0, h); P[2] = mk(0, -h); P[3] = mk(h, 0);
      two_level(P, 2, p[0], g, m);
      return 2;
    case QHBM_GATE_ZPOW:
      P[3] = one;
      two_level(P, 2, p[0], g, m);
      return 2;
    case QHBM_GATE_HPOW:  // (I - H)/2
      P[0] = mk(h - h * r, 0); P[1] = mk(-h * r, 0); P[2] = mk(-h * r, 0); P[3] = mk(h + h * r, 0);
      two_level(P, 2, p[0], g, m);
      return 2;
    case QHBM_GATE_CZPOW:
      P[15] = one;
      two_level(P, 4, p[0], g, m);
      return 4;
    case QHBM_GATE_CNOTPOW:  // |1><1| (x) (I-X)/2
      P[10] = mk(h, 0); P[11] = mk(-h, 0); P[14] = mk(-h, 0); P[15] = mk(h, 0);
      two_level(P, 4, p[0], g, m);
      return 4;
    case QHBM_GATE_SWAPPOW:  // (I - SWAP)/2
      P[5] = mk(h, 0); P[6] = mk(-h, 0); P[9] = mk(-h, 0); P[10] = mk(h, 0);
      two_level(P, 4, p[0], g, m);
      return 4;
    case QHBM_GATE_XXPOW:  // (I - XX)/2, XX = antidiagonal ones
      for (int i = 0; i < 4; ++i) { P[i * 4 + i] = mk(h, 0); P[i * 4 + (3 - i)] = mk(-h, 0); }
      two_level(P, 4, p[0], g, m);
      return 4;
    case QHBM_GATE_YYPOW:  // YY = antidiag(-1, 1, 1, -1)
      for (int i = 0; i < 4; ++i) P[i * 4 + i] = mk(h, 0);
      P[0 * 4 + 3] = mk(h, 0); P[1 * 4 + 2] = mk(-h, 0); P[2 * 4 + 1] = mk(-h, 0); P[3 * 4 + 0] = mk(h, 0);
      two_level(P, 4, p[0], g, m);
      return 4;
    case QHBM_GATE_ZZPOW:  // (I - ZZ)/2 = diag(0,1,1,0)
      P[5] = one; P[10] = one;
      two_level(P, 4, p[0], g, m);
      return 4;
    case QHBM_GATE_ISWAPPOW: {
      cd ph = expipi(p[0] * g);
      cd e = expipi(0.5 * p[0]);  // c + i s
      m[0] = ph; m[15] = ph;
      m[5] = ph * mk(e.re, 0); m[10] = m[5];
      m[6] = ph * mk(0, e.im); m[9] = m[6];
      return 4;
    }
    case QHBM_GATE_PHASEDXPOW: {  // Z^ph X^t Z^-ph
      P[0] = mk(h, 0); P[1] = mk(-h, 0); P[2] = mk(-h, 0); P[3] = mk(h, 0);
      two_level(P, 2, p[0], g, m);
      cd f = expipi(p[1]);
      m[1] = m[1] * conj(f);
      m[2] = m[2] * f;
      return 2;
    }
    case QHBM_GATE_FSIM: {
      double c = cos(p[0]), s = sin(p[0]);
      m[0] = one; m[5] = mk(c, 0); m[10] = mk(c, 0);
      m[6] = mk(0, -s); m[9] = mk(0, -s);
      m[15] = expi(-p[1]);
      return 4;
    }
    case QHBM_GATE_PHASEDISWAPPOW: {
      cd e = expipi(0.5 * p[0]);
      cd f = expipi(2.0 * p[1]);
      m[0] = one; m[15] = one;
      m[5] = mk(e.re, 0); m[10] = mk(e.re, 0);
      m[6] = mk(0, e.im) * f;
      m[9] = mk(0, e.im) * conj(f);
      return 4;
    }
    default:
      m[0] = one; m[3] = one;
      return 2;
  }
}

QHBM_HD void gate_param_values(const qhbm_gate_t& g, const float* symbols, double* p) {
  for (int k = 0; k < 3; ++k) {
    double v = 0.0;
    if (k < g.nparams) {
      v = (double)g.cnst[k];
      if (g.sym[k] >= 0) v += (double)g.scalar[k] * (double)symbols[g.sym[k]];
    }
    p[k] = v;
  }
}

QHBM_HD int gate_matrix_of(const qhbm_gate_t& g, const float* symbols, cd* m) {
  double p[3];
  gate_param_values(g, symbols, p);
  return gate_matrix(g.type, p, (double)g.gshift, m);
}

// d(matrix)/d(symbol bound to parameter k).  mode: enum qhbm_grad_mode.
QHBM_HD int gate_derivative(const qhbm_gate_t& g, const float* symbols, int k, int mode, cd* dm) {
  double p[3], q[3];
  gate_param_values(g, symbols, p);
  const double a = (double)g.scalar[k];
  cd t0[16], t1[16];
  int dim;
  if (mode == QHBM_GRAD_EXACT) {
    // 4th-order central stencil in float64, h = 1e-3 on the symbol: truncation
    // error ~1e-11 relative, i.e. exact for every tolerance in this project.
    const double h = 1e-3;
    cd t2[16], t3[16];
    for (int i = 0; i < 3; ++i) q[i] = p[i];
    q[k] = p[k] + 2 * a * h; dim = gate_matrix(g.type, q, (double)g.gshift, t0);
    q[k] = p[k] + a * h;           gate_matrix(g.type, q, (double)g.gshift, t1);
    q[k] = p[k] - a * h;           gate_matrix(g.type, q, (double)g.gshift, t2);
    q[k] = p[k] - 2 * a * h;       gate_matrix(g.type, q, (double)g.gshift, t3);
    const double s = 1.0 / (12.0 * h);
    for (int i = 0; i < dim * dim; ++i)
      dm[i] = s * (8.0 * (t1[i] - t2[i]) - (t0[i] - t3[i]));
    return dim;
  }
  const double eps = 5e-3;  // TFQ adj_util.cc _GRAD_EPS
  for (int i = 0; i < 3; ++i) q[i] = p[i];
  q[k] = p[k] + a * eps; dim = gate_matrix(g.type, q, (double)g.gshift, t0);
  q[k] = p[k] - a * eps;       gate_matrix(g.type, q, (double)g.gshift, t1);
  if (mode == QHBM_GRAD_TFQ_FD_F32) {
    const float s = 0.5f * (1.0f / 5e-3f);
    for (int i = 0; i < dim * dim; ++i) {
      float ar = (float)t0[i].re, ai = (float)t0[i].im, br = (float)t1[i].re, bi = (float)t1[i].im;
      dm[i] = mk((double)((ar - br) * s), (double)((ai - bi) * s));
    }
  } else {
    const double s = 0.5 / eps;
    for (int i = 0; i < dim * dim; ++i) dm[i] = s * (t0[i] - t1[i]);
  }
  return dim;
}

// c = a * b (dim x dim)
QHBM_HD void matmul(const cd* a, const cd* b, int dim, cd* c) {
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < dim; ++j) {
      cd s = mk(0, 0);
      for (int k = 0; k < dim; ++k) s = s + a[i * dim + k] * b[k * dim + j];
      c[i * dim + j] = s;
    }
}
QHBM_HD void dagger(const cd* a, int dim, cd* c) {
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < dim; ++j) c[i * dim + j] = conj(a[j * dim + i]);
}
// Exchange the roles of the two qubits of a 4x4 matrix (index bit swap).
QHBM_HD void swap_qubits(const cd* a, cd* c) {
  const int perm[4] = {0, 2, 1, 3};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) c[i * 4 + j] = a[perm[i] * 4 + perm[j]];
}

}  // namespace qhbm
