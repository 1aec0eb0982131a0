// measure.cu -- sm_100a kernels + C-ABI for the shot-based side of quantum inference
// (reference qhbmlib/inference/qnn.py:142-292, SampledQuantumInference):
//   * qhbm_sample_states   bitstring shots from |psi_u|^2 of many final states (tfq.layers.Sample)
//   * qhbm_binomial_shots  shot-noise of a +-1 valued measurement whose exact mean is known
//                          (tfq.layers.SampledExpectation measures every Pauli term with its own
//                          `repetitions` shots, so the term estimate is (2 Binomial(R, (1+<P>)/2) - R)/R)
// Both are bandwidth/latency-bound scan-and-search kernels: no tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <stdexcept>

#include "../../include/qhbm_b200.h"
#include "common.h"
#include "philox.h"

namespace qhbm {

constexpr int kShotThreads = 256;

// grid = (n_states, sample slices).  Each CTA rebuilds the 256-entry coarse CDF of its state
// (warp-cooperative, coalesced), then every thread resolves its shots: binary search in shared
// memory, linear scan inside one chunk of the state (L1/L2 resident after the first pass).
__global__ void __launch_bounds__(kShotThreads)
    shot_sample_kernel(const float2* __restrict__ states, int n_qubits, const int64_t* __restrict__ offsets,
                       uint64_t seed0, uint64_t seed1, uint64_t* __restrict__ out) {
  __shared__ double s_pre[kShotThreads + 1];
  const int64_t u = blockIdx.x;
  const int64_t dim = (int64_t)1 << n_qubits;
  const float2* __restrict__ psi = states + u * dim;
  const int64_t first = offsets[u], count = offsets[u + 1] - first;
  if (count <= 0) return;
  const int n_chunks = (int)min((int64_t)kShotThreads, dim);
  const int64_t per = dim / n_chunks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp * 32; c < warp * 32 + 32; ++c) {
    double acc = 0.0;
    if (c < n_chunks) {
      const float2* __restrict__ base = psi + (int64_t)c * per;
      for (int64_t i = lane; i < per; i += 32) {
        const float2 a = base[i];
        acc += (double)a.x * a.x + (double)a.y * a.y;
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_pre[c + 1] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = 0.0;
    s_pre[0] = 0.0;
    for (int c = 1; c <= kShotThreads; ++c) {
      run += s_pre[c];
      s_pre[c] = run;
    }
  }
  __syncthreads();
  const double total = s_pre[n_chunks];
  const Philox ph = make_philox(seed0, seed1);
  for (int64_t r = (int64_t)blockIdx.y * kShotThreads + threadIdx.x; r < count;
       r += (int64_t)gridDim.y * kShotThreads) {
    const uint4 rnd = ph((uint64_t)(first + r), 0x53484F54u);
    const double target = u01_53(rnd.x, rnd.y) * total;
    int lo = 0, hi = n_chunks - 1;  // largest chunk with s_pre[chunk] <= target
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_pre[mid] <= target) lo = mid; else hi = mid - 1;
    }
    double acc = s_pre[lo];
    const int64_t i0 = (int64_t)lo * per, i1 = i0 + per;
    int64_t pick = -1, last_pos = -1;
    for (int64_t i = i0; i < i1; ++i) {
      const float2 a = psi[i];
      const double w = (double)a.x * a.x + (double)a.y * a.y;
      if (w > 0.0) last_pos = i;
      acc += w;
      if (target < acc) { pick = i; break; }
    }
    if (pick < 0) pick = last_pos >= 0 ? last_pos : i0;  // rounding pushed the target past the chunk
    out[first + r] = (uint64_t)pick;
  }
}

// Stirling tail log(k!) - [(k+1/2) log(k+1) - (k+1) + log(2 pi)/2].
__device__ __forceinline__ double stirling_tail(double k) {
  const double tab[10] = {0.0810614667953272, 0.0413406959554092, 0.0276779256849983, 0.02079067210376509,
                          0.0166446911898211, 0.0138761288230707, 0.0118967099458917, 0.0104112652619720,
                          0.00925546218271273, 0.00833056343336287};
  if (k <= 9.0) return tab[(int)k];
  const double kp1 = k + 1.0, kp1sq = kp1 * kp1;
  return (1.0 / 12 - (1.0 / 360 - 1.0 / 1260 / kp1sq) / kp1sq) / kp1;
}

struct Draws {
  Philox ph;
  uint64_t counter;
  uint32_t call;
  uint4 buf;
  int left;
  __device__ __forceinline__ double next() {
    if (left == 0) {
      buf = ph(counter, 0x42494E00u + call++);
      left = 2;
    }
    --left;
    return left == 1 ? u01_53(buf.x, buf.y) : u01_53(buf.z, buf.w);
  }
};

// Exact Binomial(n, p) for p <= 1/2: sum of geometric gaps for small means, Hormann's transformed
// rejection with squeeze (BTRS) otherwise.
__device__ double binomial_draw(Draws& d, double n, double p) {
  if (p <= 0.0 || n <= 0.0) return 0.0;
  if (n * p < 10.0) {
    const double lq = log1p(-p);
    double successes = 0.0, pos = 0.0;
    for (;;) {
      const double uu = fmax(d.next(), 1e-300);
      pos += ceil(log(uu) / lq);
      if (pos > n) return successes;
      successes += 1.0;
    }
  }
  const double spq = sqrt(n * p * (1.0 - p));
  const double b = 1.15 + 2.53 * spq;
  const double a = -0.0873 + 0.0248 * b + 0.01 * p;
  const double c = n * p + 0.5;
  const double v_r = 0.92 - 4.2 / b;
  const double r = p / (1.0 - p);
  const double alpha = (2.83 + 5.1 / b) * spq;
  const double m = floor((n + 1.0) * p);
  for (;;) {
    const double uu = d.next() - 0.5;
    double v = d.next();
    const double us = 0.5 - fabs(uu);
    const double k = floor((2.0 * a / us + b) * uu + c);
    if (us >= 0.07 && v <= v_r) return k;
    if (k < 0.0 || k > n) continue;
    v = log(v * alpha / (a / (us * us) + b));
    const double bound = (m + 0.5) * log((m + 1.0) / (r * (n - m + 1.0))) +
                         (n + 1.0) * log((n - m + 1.0) / (n - k + 1.0)) +
                         (k + 0.5) * log(r * (n - k + 1.0) / (k + 1.0)) + stirling_tail(m) +
                         stirling_tail(n - m) - stirling_tail(k) - stirling_tail(n - k);
    if (v <= bound) return k;
  }
}

__global__ void binomial_shots_kernel(const float* __restrict__ exact, int64_t n, int64_t shots, uint64_t seed0,
                                      uint64_t seed1, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double e = fmin(1.0, fmax(-1.0, (double)exact[i]));
  const double p_plus = 0.5 * (1.0 + e);
  Draws d;
  d.ph = make_philox(seed0, seed1);
  d.counter = (uint64_t)i;
  d.call = 0;
  d.left = 0;
  const bool flip = p_plus > 0.5;
  const double k = binomial_draw(d, (double)shots, flip ? 1.0 - p_plus : p_plus);
  const double plus = flip ? (double)shots - k : k;
  out[i] = (float)((2.0 * plus - (double)shots) / (double)shots);
}

}  // namespace qhbm

using namespace qhbm;

extern "C" {

int qhbm_sample_states(const float* d_states, int64_t n_states, int32_t n_qubits, const int64_t* d_offsets,
                       int64_t total_samples, uint64_t seed0, uint64_t seed1, uint64_t* d_out, void* stream) {
  return guarded([&] {
    if (n_qubits < 1 || n_qubits > 30) throw std::runtime_error("n_qubits must be in [1, 30]");
    if (n_states < 0 || total_samples < 0) throw std::runtime_error("negative size");
    if (n_states == 0 || total_samples == 0) return;
    if (n_states > 65535ll * 32768ll) throw std::runtime_error("too many states");
    if (!d_states || !d_offsets || !d_out) throw std::runtime_error("null pointer");
    const int64_t mean = (total_samples + n_states - 1) / n_states;
    const int slices = (int)std::min<int64_t>(64, std::max<int64_t>(1, mean / (4 * kShotThreads)));
    dim3 grid((unsigned)n_states, (unsigned)slices);
    shot_sample_kernel<<<grid, kShotThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(d_states), n_qubits, d_offsets, seed0, seed1, d_out);
    QHBM_CUDA(cudaGetLastError());
  });
}

int qhbm_binomial_shots(const float* d_exact, int64_t n, int64_t shots, uint64_t seed0, uint64_t seed1,
                        float* d_out, void* stream) {
  return guarded([&] {
    if (shots < 1) throw std::runtime_error("shots must be >= 1");
    if (n < 0) throw std::runtime_error("negative size");
    if (n == 0) return;
    if (!d_exact || !d_out) throw std::runtime_error("null pointer");
    binomial_shots_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_exact, n, shots, seed0,
                                                                                      seed1, d_out);
    QHBM_CUDA(cudaGetLastError());
  });
}

}  // extern "C"
