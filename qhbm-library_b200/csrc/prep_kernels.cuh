// prep_kernels.cuh -- the two small kernels around the sweeps of a call: coefficient preparation (symbols ->
// gate matrices, gradient matrices, phase tables; also clears the call's accumulators) and the float64 ->
// float32 write-out.  Included by sim_kernels.cuh inside namespace qhbm (not by sim_lean.cu, which only
// instantiates the sweep kernels).
#pragma once
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

// The cooperative jobs (whole CTA): constants and diagonal phase tables of one symbol row.
__device__ __forceinline__ void prep_cooperative(const PrepJob& job, const int32_t* __restrict__ list,
                                                 const qhbm_gate_t* __restrict__ gates,
                                                 const float* __restrict__ symbols, float* __restrict__ out,
                                                 const int tid) {
  if (job.kind == PJ_CONST) {
    for (int i = tid; i < job.list_len; i += kPrepThreads) out[i] = __int_as_float(list[i]);
    return;
  }
  {
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
      if (v >= entries) continue;
      if (job.b) {  // register phase table: (re, im, -im, im), the form the packed complex multiply reads
        out[4 * v + 0] = (float)acc[e].re;
        out[4 * v + 1] = (float)acc[e].im;
        out[4 * v + 2] = -(float)acc[e].im;
        out[4 * v + 3] = (float)acc[e].im;
      } else {
        write_c(out, v, acc[e]);
      }
    }
    return;
  }
}

// Symbol rows per CTA when the call has one row of symbol values per state (qhbm_expectation_*_rows): the
// per-thread jobs then run one ROW per thread, the cooperative ones loop over the CTA's rows.
constexpr int kPrepRows = kPrepThreads;

// CTAs [0, n_jobs) in x run one coefficient job each; the CTAs after them clear the float64 accumulators of
// the call (so a call needs no separate memsets).  n_rows == 1: all states share the symbol values; thread 0
// runs the per-thread jobs.  n_rows > 1: gridDim.y = ceil(n_rows / kPrepRows); row r reads
// symbols + r * sym_stride and writes its table at coef + r * coef_stride.
__global__ void __launch_bounds__(kPrepThreads) prep_kernel(const PrepJob* __restrict__ jobs, int n_jobs,
                                                            const int32_t* __restrict__ lists,
                                                            const qhbm_gate_t* __restrict__ gates,
                                                            const float* __restrict__ symbols,
                                                            float* __restrict__ coef, int mode,
                                                            double* __restrict__ zero_a, int64_t n_a,
                                                            double* __restrict__ zero_b, int64_t n_b,
                                                            int n_rows, uint32_t sym_stride, uint32_t coef_stride) {
  if ((int)blockIdx.x >= n_jobs) {
    if (blockIdx.y != 0) return;
    const int64_t stride = (int64_t)(gridDim.x - n_jobs) * kPrepThreads;
    for (int64_t i = (int64_t)(blockIdx.x - n_jobs) * kPrepThreads + threadIdx.x; i < n_a + n_b; i += stride) {
      if (i < n_a) zero_a[i] = 0.0;
      else zero_b[i - n_a] = 0.0;
    }
    return;
  }
  const PrepJob job = jobs[blockIdx.x];
  const int32_t* list = lists + job.list_off;
  const int tid = threadIdx.x;
  const bool multi = n_rows > 1;
  const int row0 = multi ? (int)blockIdx.y * kPrepRows : 0;
  if (job.kind == PJ_CONST || job.kind == PJ_DTAB) {
    const int row1 = multi ? min(n_rows, row0 + kPrepRows) : 1;
    for (int row = row0; row < row1; ++row)
      prep_cooperative(job, list, gates, symbols + (size_t)row * sym_stride, coef + (size_t)row * coef_stride + job.out, tid);
    return;
  }
  const int row = multi ? row0 + tid : 0;
  if (multi ? row >= n_rows : tid != 0) return;
  symbols += (size_t)row * sym_stride;
  float* out = coef + (size_t)row * coef_stride + job.out;
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
    case PJ_GDIAG:
    case PJ_KAPPA: {
      const qhbm_gate_t g = gates[list[0]];
      const int dim = gate_matrix_of(g, symbols, m);
      gate_derivative(g, symbols, job.c, mode, t);
      dagger(m, dim, w);
      matmul(t, w, dim, m);  // M = dG G^dagger
      if (dim == 4 && job.b && job.kind != PJ_KAPPA) { swap_qubits(m, t); for (int k = 0; k < 16; ++k) m[k] = t[k]; }
      if (job.kind == PJ_GDIAG) {
        for (int k = 0; k < 4; ++k) write_c(out, k, k < dim ? m[k * dim + k] : mk(0, 0));
      } else if (job.kind == PJ_KAPPA) {
        // M = i c0 I - i (kappa/2) A:  X: M01 = -i kappa/2;  Y: M01 = -kappa/2
        out[0] = (float)(job.b == 0 ? -2.0 * m[1].im : -2.0 * m[1].re);
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
    case PJ_ROT: {
      if (job.list_len == 0) {  // identity rotation filling an unused position of OP_XROTF
        out[0] = 1.f; out[1] = 0.f; out[2] = 0.f; out[3] = 0.f;
        break;
      }
      double p[3];
      gate_param_values(gates[list[0]], symbols, p);
      const cd e = expipi(0.5 * p[0]);
      out[0] = (float)e.re;
      out[1] = (float)(job.a ? -e.im : e.im);
    } break;
    case PJ_ROTF: {
      const int K = job.d;
      double c[kMaxRegQubits], sn[kMaxRegQubits], kap[kMaxRegQubits];
      bool fast = true;
      for (int P = 0; P < K; ++P) {
        c[P] = 1.0; sn[P] = 0.0; kap[P] = 0.0;
        if (list[P] < 0) continue;
        const qhbm_gate_t g = gates[list[P]];
        double pv[3];
        gate_param_values(g, symbols, pv);
        const cd e = expipi(0.5 * pv[0]);
        c[P] = e.re;
        sn[P] = job.a ? -e.im : e.im;
        if (job.b & (1 << P)) {
          const int dim = gate_matrix_of(g, symbols, m);
          gate_derivative(g, symbols, 0, mode, t);
          dagger(m, dim, w);
          matmul(t, w, dim, m);  // M = dG G^dagger = i c0 I - i (kappa/2) X
          kap[P] = -2.0 * m[1].im;
        }
        if (P < K - 1 && fabs(c[P]) < 0.05) fast = false;
      }
      double scale = 1.0;  // product of the cosines of the unnormalised positions so far
      for (int P = 0; P < K; ++P) {
        if (fast && P < K - 1) {
          out[4 * P + 0] = (float)(sn[P] / c[P]);
          out[4 * P + 1] = 0.f;
          out[4 * P + 2] = (float)(kap[P] * scale * scale);
          scale *= c[P];
        } else if (fast) {
          out[4 * P + 0] = (float)(c[P] * scale);
          out[4 * P + 1] = (float)(sn[P] * scale);
          out[4 * P + 2] = (float)(kap[P] * scale * scale);
        } else {
          out[4 * P + 0] = (float)c[P];
          out[4 * P + 1] = (float)sn[P];
          out[4 * P + 2] = (float)kap[P];
        }
        out[4 * P + 3] = fast ? 1.f : 0.f;
      }
    } break;
    case PJ_PHASE: {
      cd acc = mk(1, 0);
      for (int i = 0; i < job.list_len; ++i) {
        const qhbm_gate_t g = gates[list[i]];
        double p[3];
        gate_param_values(g, symbols, p);
        acc = acc * expipi(p[0] * ((double)g.gshift + 0.5));
      }
      write_c(out, 0, acc);
    } break;
    default: break;
  }
}

// float64 accumulators -> the caller's float32 outputs (expectations and, if present, gradients)
__global__ void finalize_kernel(const double* __restrict__ src_a, float* __restrict__ dst_a, int64_t n_a,
                                const double* __restrict__ src_b, float* __restrict__ dst_b, int64_t n_b) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_a) dst_a[i] = (float)src_a[i];
  else if (i < n_a + n_b) dst_b[i - n_a] = (float)src_b[i - n_a];
}
