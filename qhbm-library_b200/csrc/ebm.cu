// ebm.cu -- sm_100a kernels + C-ABI for the classical (EBM) side of the hot path:
// bitstring packing, first-occurrence unique-with-counts, energy evaluation, the
// exhaustive 2^n logits / logsumexp / entropy sweep, categorical and Bernoulli
// sampling, count-weighted reductions.  Reference interfaces replaced are cited in
// include/qhbm_b200.h.  These are HBM/latency-bound integer and fp32 kernels: no
// tensor cores (energies feed exp() and must keep fp32 accuracy).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <utility>
#include <stdexcept>
#include <vector>

#include "../../include/qhbm_b200.h"
#include "common.h"
#include "philox.h"

namespace qhbm {

// ------------------------------------------------------------------ pack / unpack
struct ShiftTable {
  int8_t s[64];
};

__global__ void pack_kernel(const int8_t* __restrict__ bits, int64_t n_rows, int n_bits, ShiftTable st,
                            uint64_t* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int8_t* row = bits + i * n_bits;
  uint64_t k = 0;
  for (int j = 0; j < n_bits; ++j) k |= (uint64_t)(row[j] & 1) << st.s[j];
  keys[i] = k;
}
__global__ void unpack_kernel(const uint64_t* __restrict__ keys, int64_t n_rows, int n_bits, ShiftTable st,
                              int8_t* __restrict__ bits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * n_bits) return;
  const int64_t r = i / n_bits;
  const int j = (int)(i - r * n_bits);
  bits[i] = (int8_t)((keys[r] >> st.s[j]) & 1);
}

// ------------------------------------------------------------------ unique with counts
constexpr uint64_t kEmptyKey = ~0ull;
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
struct UniqueWs {
  unsigned long long* tkeys;  // [cap]
  int32_t* tfirst;            // [cap] first row of the key, later its rank
  int32_t* tcount;            // [cap]
  int32_t* slot;              // [N]
  int32_t* rank;              // [N] exclusive scan of first-occurrence flags
  int32_t* bsum;              // [nblocks + 1]
  uint64_t cap;
};
__global__ void uq_init(UniqueWs w) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w.cap) w.tkeys[i] = kEmptyKey;
  if (i <= w.cap) { w.tfirst[i] = 0x7fffffff; w.tcount[i] = 0; }  // entry `cap`: the key equal to kEmptyKey
}
__global__ void uq_insert(const uint64_t* __restrict__ keys, int64_t n, UniqueWs w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i];
  uint64_t h = mix64(k) & (w.cap - 1);
  if (k == kEmptyKey) {
    h = w.cap;  // the one key that cannot live in the table (it marks empty slots) has its own entry
  } else {
    while (true) {
      const unsigned long long old = atomicCAS(&w.tkeys[h], kEmptyKey, k);
      if (old == kEmptyKey || old == k) break;
      h = (h + 1) & (w.cap - 1);
    }
  }
  atomicMin(&w.tfirst[h], (int32_t)i);
  atomicAdd(&w.tcount[h], 1);
  w.slot[i] = (int32_t)h;
}
constexpr int kScanBlock = 1024;
// flags -> per-block exclusive scan (in rank) + block totals
__global__ void __launch_bounds__(kScanBlock) uq_flag_scan(int64_t n, UniqueWs w) {
  __shared__ int32_t s_w[32];
  const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int32_t f = 0;
  if (i < n) f = (w.tfirst[w.slot[i]] == (int32_t)i) ? 1 : 0;
  int32_t v = f;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) s_w[wid] = v;
  __syncthreads();
  if (wid == 0) {
    int32_t x = s_w[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += t;
    }
    s_w[lane] = x;
  }
  __syncthreads();
  const int32_t incl = v + (wid ? s_w[wid - 1] : 0);
  if (i < n) w.rank[i] = incl - f;
  if (threadIdx.x == kScanBlock - 1) w.bsum[blockIdx.x] = incl;
}
// exclusive scan of the block totals by ONE block (nblocks <= ~1e6 is fine)
__global__ void __launch_bounds__(1024) uq_scan_blocks(int64_t nblocks, UniqueWs w, int64_t* n_unique) {
  __shared__ int32_t s_w[32];
  __shared__ int32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t b0 = 0; b0 < nblocks; b0 += 1024) {
    const int64_t i = b0 + threadIdx.x;
    const int32_t f = i < nblocks ? w.bsum[i] : 0;
    int32_t v = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) s_w[wid] = v;
    __syncthreads();
    if (wid == 0) {
      int32_t x = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
      }
      s_w[lane] = x;
    }
    __syncthreads();
    const int32_t carry = s_carry;
    const int32_t incl = v + (wid ? s_w[wid - 1] : 0) + carry;
    if (i < nblocks) w.bsum[i] = incl - f;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_unique = s_carry;
}
__global__ void uq_emit(const uint64_t* __restrict__ keys, int64_t n, UniqueWs w, uint64_t* __restrict__ uniq,
                        int32_t* __restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t s = w.slot[i];
  const int32_t r = w.rank[i] + w.bsum[i / kScanBlock];
  w.rank[i] = r;
  if (w.tfirst[s] == (int32_t)i) {
    uniq[r] = keys[i];
    count[r] = w.tcount[s];
  }
}
// second phase: slot -> rank of its first occurrence (first rows hold their own rank)
__global__ void uq_index(int64_t n, UniqueWs w, int32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  idx[i] = w.rank[w.tfirst[w.slot[i]]];
}

__global__ void segsum_kernel(const float* __restrict__ vals, const int32_t* __restrict__ idx, int64_t n_rows,
                              int width, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * width) return;
  const int64_t r = i / width;
  const int c = (int)(i - r * width);
  atomicAdd(&out[(int64_t)idx[r] * width + c], vals[i]);
}

constexpr int kWsumStrip = 256;
__global__ void wsum_kernel(const int32_t* __restrict__ counts, const float* __restrict__ vals, int64_t n_rows,
                            int width, double* __restrict__ out) {
  const int64_t r0 = (int64_t)blockIdx.x * kWsumStrip;
  const int64_t r1 = min(n_rows, r0 + kWsumStrip);
  for (int c = threadIdx.x; c <= width; c += blockDim.x) {
    double s = 0.0;
    if (c < width) {
      for (int64_t r = r0; r < r1; ++r) s += (double)counts[r] * (double)vals[r * width + c];
    } else {
      for (int64_t r = r0; r < r1; ++r) s += (double)counts[r];
    }
    atomicAdd(&out[c], s);
  }
}

// ------------------------------------------------------------------ score-function gradient (f1)
// EnergyInference._expectation's backward for energies E = sum_t theta_t f_t(x) with parity features
// f_t(x) = (-1)^{parity(x & mask_t)} (Bernoulli: single-bit masks, KOBE: Z-strings; reference
// ebm.py:282-325 with the Jacobian of energy_utils.py:97-110):
//   d/dtheta_t = E[c] E[f_t] - E[c f_t],  c_u = sum_j upstream_j values[u, j],  E[c] = upstream . average.
// Kernel 1: c_u for every row (+ E[c]).  Kernel 2: one CTA per term, float64 accumulation.
__global__ void score_c_kernel(const float* __restrict__ vals, int64_t n_rows, int width,
                               const float* __restrict__ upstream, const float* __restrict__ avg,
                               float* __restrict__ c_out, float* __restrict__ s0_out) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u == 0) {
    double s0 = 0.0;
    for (int j = 0; j < width; ++j) s0 += (double)upstream[j] * (double)avg[j];
    *s0_out = (float)s0;
  }
  if (u >= n_rows) return;
  float c = 0.f;
  for (int j = 0; j < width; ++j) c = fmaf(upstream[j], vals[u * width + j], c);
  c_out[u] = c;
}
__global__ void __launch_bounds__(256) score_term_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ counts,
                                                         int64_t n_rows, const float* __restrict__ c,
                                                         const float* __restrict__ s0, const int32_t* __restrict__ masks,
                                                         const double* __restrict__ total, float scale,
                                                         float* __restrict__ out) {
  __shared__ double s_w[8];
  const uint64_t mask = (uint64_t)(uint32_t)masks[blockIdx.x];
  const float e_c = *s0;
  double acc = 0.0;
  for (int64_t u = threadIdx.x; u < n_rows; u += blockDim.x) {
    const float f = (__popcll(keys[u] & mask) & 1) ? -1.f : 1.f;
    acc += (double)counts[u] * (double)(f * (e_c - c[u]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += s_w[k];
    out[blockIdx.x] = (float)((double)scale * t / *total);
  }
}

// ------------------------------------------------------------------ energies
constexpr int kMaxWidth = 64;
struct EnergyArgs {
  qhbm_energy_desc_t d;
};

// Dense stack on the raw bits; one row per thread, activations in registers, weights in
// shared memory read as broadcast float4 (4 output neurons per load).
__device__ __forceinline__ float mlp_energy(const EnergyArgs& ea, const float* s_w, uint64_t key) {
  const qhbm_energy_desc_t& d = ea.d;
  float x[kMaxWidth], y[kMaxWidth];
  const int n = d.n_bits;
#pragma unroll
  for (int j = 0; j < kMaxWidth; ++j) x[j] = j < n ? (float)((key >> (n - 1 - j)) & 1) : 0.f;
  int off = 0;
  for (int l = 0; l < d.n_layers; ++l) {
    const int in = d.widths[l], out = d.widths[l + 1];
    const int out4 = (out + 3) & ~3;  // rows padded to 4 outputs in smem
    const float* W = s_w + off;
    const float* Bv = W + in * out4;
#pragma unroll
    for (int o = 0; o < kMaxWidth; o += 4) {
      if (o < out) {
        const float4 b = *reinterpret_cast<const float4*>(Bv + o);
        y[o] = b.x; y[o + 1] = b.y; y[o + 2] = b.z; y[o + 3] = b.w;
      }
    }
#pragma unroll
    for (int i = 0; i < kMaxWidth; ++i) {
      if (i < in) {
        const float xi = x[i];
#pragma unroll
        for (int o = 0; o < kMaxWidth; o += 4) {
          if (o < out) {
            const float4 w4 = *reinterpret_cast<const float4*>(W + i * out4 + o);
            y[o] = fmaf(xi, w4.x, y[o]);
            y[o + 1] = fmaf(xi, w4.y, y[o + 1]);
            y[o + 2] = fmaf(xi, w4.z, y[o + 2]);
            y[o + 3] = fmaf(xi, w4.w, y[o + 3]);
          }
        }
      }
    }
    const int act = d.act[l];
#pragma unroll
    for (int o = 0; o < kMaxWidth; ++o) {
      float v = y[o];
      if (act == 1) v = tanhf(v);
      else if (act == 2) v = fmaxf(v, 0.f);
      x[o] = o < out ? v : 0.f;
    }
    off += in * out4 + out4;
  }
  return x[0];
}

__device__ __forceinline__ float parity_energy(const EnergyArgs& ea, const float* s_theta, const uint32_t* s_mask,
                                               uint64_t key) {
  const uint32_t k = (uint32_t)key;
  float e = 0.f;
  for (int t = 0; t < ea.d.n_terms; ++t) {
    const uint32_t sgn = (uint32_t)(__popc(k & s_mask[t]) & 1) << 31;
    e += __uint_as_float(__float_as_uint(s_theta[t]) ^ sgn);
  }
  return e;
}

// Stages the energy parameters in shared memory.  Returns the float count used.
__device__ __forceinline__ void stage_energy(const EnergyArgs& ea, float* s_f) {
  const qhbm_energy_desc_t& d = ea.d;
  if (d.kind == QHBM_ENERGY_MLP) {
    int off = 0;
    for (int l = 0; l < d.n_layers; ++l) {
      const int in = d.widths[l], out = d.widths[l + 1];
      const int out4 = (out + 3) & ~3;
      for (int i = threadIdx.x; i < in * out4; i += blockDim.x) {
        const int r = i / out4, c = i - r * out4;
        s_f[off + i] = c < out ? d.d_weights[l][r * out + c] : 0.f;
      }
      for (int i = threadIdx.x; i < out4; i += blockDim.x) s_f[off + in * out4 + i] = i < out ? d.d_bias[l][i] : 0.f;
      off += in * out4 + out4;
    }
  } else {
    uint32_t* s_m = reinterpret_cast<uint32_t*>(s_f + d.n_terms);
    for (int i = threadIdx.x; i < d.n_terms; i += blockDim.x) {
      s_f[i] = d.d_theta[i];
      s_m[i] = d.d_masks[i];
    }
  }
  __syncthreads();
}
__device__ __forceinline__ float eval_energy(const EnergyArgs& ea, const float* s_f, uint64_t key) {
  if (ea.d.kind == QHBM_ENERGY_MLP) return mlp_energy(ea, s_f, key);
  return parity_energy(ea, s_f, reinterpret_cast<const uint32_t*>(s_f + ea.d.n_terms), key);
}

size_t energy_smem_bytes(const qhbm_energy_desc_t& d) {
  if (d.kind == QHBM_ENERGY_MLP) {
    size_t f = 0;
    for (int l = 0; l < d.n_layers; ++l) {
      const int out4 = (d.widths[l + 1] + 3) & ~3;
      f += (size_t)d.widths[l] * out4 + out4;
    }
    return f * 4;
  }
  return (size_t)d.n_terms * 8;
}

void validate_energy(const qhbm_energy_desc_t* e) {
  if (!e) throw std::runtime_error("null energy descriptor");
  if (e->n_bits < 1 || e->n_bits > 62) throw std::runtime_error("energy n_bits out of range");
  if (e->kind == QHBM_ENERGY_MLP) {
    if (e->n_layers < 1 || e->n_layers > 8) throw std::runtime_error("MLP energy: 1..8 layers supported");
    if (e->widths[0] != e->n_bits) throw std::runtime_error("MLP energy: widths[0] must equal n_bits");
    for (int l = 0; l <= e->n_layers; ++l)
      if (e->widths[l] < 1 || e->widths[l] > kMaxWidth) throw std::runtime_error("MLP energy: widths must be in [1, 64]");
    if (e->widths[e->n_layers] != 1) throw std::runtime_error("MLP energy: last layer must have one output");
    for (int l = 0; l < e->n_layers; ++l)
      if (!e->d_weights[l] || !e->d_bias[l]) throw std::runtime_error("MLP energy: null layer pointer");
  } else if (e->kind == QHBM_ENERGY_BERNOULLI || e->kind == QHBM_ENERGY_KOBE) {
    if (e->n_bits > 32) throw std::runtime_error("parity energies support up to 32 bits");
    if (e->n_terms < 0 || (e->n_terms > 0 && (!e->d_masks || !e->d_theta))) throw std::runtime_error("bad parity energy tables");
  } else {
    throw std::runtime_error("unknown energy kind");
  }
  if (energy_smem_bytes(*e) > 200 * 1024) throw std::runtime_error("energy parameters do not fit shared memory");
}

constexpr int kEnergyThreads = 256;

__global__ void __launch_bounds__(kEnergyThreads) energy_rows_kernel(const __grid_constant__ EnergyArgs ea,
                                                                     const uint64_t* __restrict__ keys,
                                                                     int64_t n_rows, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_f[];
  stage_energy(ea, s_f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = eval_energy(ea, s_f, keys[i]);
}

// online (max, sum exp, sum exp*l) triple
struct Stat {
  double m, s, t;
};
__device__ __forceinline__ Stat stat_merge(Stat a, Stat b) {
  if (b.s == 0.0) return a;
  if (a.s == 0.0) return b;
  Stat r;
  r.m = fmax(a.m, b.m);
  const double fa = exp(a.m - r.m), fb = exp(b.m - r.m);
  r.s = a.s * fa + b.s * fb;
  r.t = a.t * fa + b.t * fb;
  return r;
}
// one more logit: a single exp instead of the two of a general merge
__device__ __forceinline__ void stat_add(Stat& a, double l) {
  if (a.s == 0.0) { a.m = l; a.s = 1.0; a.t = l; return; }
  if (l <= a.m) {
    const double w = exp(l - a.m);
    a.s += w;
    a.t = fma(w, l, a.t);
  } else {
    const double f = exp(a.m - l);
    a.s = fma(a.s, f, 1.0);
    a.t = fma(a.t, f, l);
    a.m = l;
  }
}
__device__ __forceinline__ Stat stat_warp(Stat v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Stat w;
    w.m = __shfl_xor_sync(0xffffffffu, v.m, o);
    w.s = __shfl_xor_sync(0xffffffffu, v.s, o);
    w.t = __shfl_xor_sync(0xffffffffu, v.t, o);
    v = stat_merge(v, w);
  }
  return v;
}

__global__ void __launch_bounds__(kEnergyThreads) ebm_sweep_kernel(const __grid_constant__ EnergyArgs ea, uint64_t lo,
                                                                   uint64_t hi, float* __restrict__ logits,
                                                                   Stat* __restrict__ partial) {
  extern __shared__ __align__(16) float s_f[];
  __shared__ Stat s_st[kEnergyThreads / 32];
  stage_energy(ea, s_f);
  Stat acc;
  acc.m = 0.0; acc.s = 0.0; acc.t = 0.0;
  for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * blockDim.x) {
    const float l = -eval_energy(ea, s_f, i);
    if (logits) logits[i - lo] = l;
    stat_add(acc, (double)l);
  }
  acc = stat_warp(acc);
  if ((threadIdx.x & 31) == 0) s_st[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    Stat v;
    v.m = 0.0; v.s = 0.0; v.t = 0.0;
    if (threadIdx.x < kEnergyThreads / 32) v = s_st[threadIdx.x];
    v = stat_warp(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}
// ------------------------------------------------------------------ dense-stack sweep, register tiled
// The dense stack over 2^n rows is a chain of small GEMMs ([rows x in] x [in x out], widths <= 64) that must
// stay in fp32 FMA (energies feed exp()).  One CTA of 128 threads pushes tiles of 128 consecutive rows
// through all layers with the activations feature-major in ONE shared-memory buffer ([feature][row]);
// each thread owns an 8-row x 8-output micro-tile (rows tx*4..+3 and 64+tx*4..+3), so four LDS.128
// feed 64 FMAs.  The first layer never multiplies: its inputs are the bits of the row index, so a
// thread adds the weight rows of the set bits -- once per tile for the bits its rows share, and from
// three preloaded weight rows for the bits (0, 1, 6) that distinguish its rows.
constexpr int kMlpRows = 128;
constexpr int kMlpThreads = 128;

__device__ __forceinline__ float fast_tanh(float x) {
  x = fminf(fmaxf(x, -9.f), 9.f);
  const float e = __expf(2.f * x);
  return __fdividef(e - 1.f, e + 1.f);
}
__device__ __forceinline__ float mlp_act(float y, int act) {
  if (act == 1) return fast_tanh(y);
  if (act == 2) return fmaxf(y, 0.f);
  return y;
}

// Padded output width of layer l in shared memory: multiples of 8 for the register-tiled layers, unpadded
// for the last layer (one output per row; the 2 KB this saves is what lets a fourth CTA fit on the SM).
__host__ __device__ inline int mlp_out_pad(const qhbm_energy_desc_t& d, int l) {
  const int out = d.widths[l + 1];
  return l == d.n_layers - 1 ? out : ((out + 7) & ~7);
}
size_t mlp_sweep_weight_floats(const qhbm_energy_desc_t& d) {
  size_t f = 0;
  for (int l = 0; l < d.n_layers; ++l) {
    const int out8 = mlp_out_pad(d, l);
    f += (size_t)d.widths[l] * out8 + out8;
  }
  return f;
}

__global__ void __launch_bounds__(kMlpThreads) ebm_mlp_sweep_kernel(const __grid_constant__ EnergyArgs ea, uint64_t lo,
                                                                    uint64_t hi, float* __restrict__ logits,
                                                                    Stat* __restrict__ partial) {
  extern __shared__ __align__(16) float s_f[];
  __shared__ Stat s_st[kMlpThreads / 32];
  const qhbm_energy_desc_t& d = ea.d;
  const int tid = threadIdx.x;
  // weights [in][out8] + bias[out8] per layer (outputs padded to 8 with zeros)
  int wfloats = 0;
  for (int l = 0; l < d.n_layers; ++l) {
    const int in = d.widths[l], out = d.widths[l + 1];
    const int out8 = mlp_out_pad(d, l);
    for (int i = tid; i < in * out8; i += kMlpThreads) {
      const int r = i / out8, c = i - r * out8;
      s_f[wfloats + i] = c < out ? d.d_weights[l][r * out + c] : 0.f;
    }
    for (int i = tid; i < out8; i += kMlpThreads) s_f[wfloats + in * out8 + i] = i < out ? d.d_bias[l][i] : 0.f;
    wfloats += in * out8 + out8;
  }
  float* buf = s_f + ((wfloats + 3) & ~3);
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;  // rows tx*4.. (+64), outputs ty*8..
  const int n = d.n_bits;
  const uint64_t t_first = lo / kMlpRows, t_last = (hi - 1) / kMlpRows;  // tiles of the ABSOLUTE row index
  Stat acc_st;
  acc_st.m = 0.0; acc_st.s = 0.0; acc_st.t = 0.0;
  for (uint64_t tile = t_first + blockIdx.x; tile <= t_last; tile += gridDim.x) {
    const uint64_t row0 = tile * kMlpRows;
    __syncthreads();  // previous tile fully consumed
    // ---- layer 0: sums of weight rows selected by the bits of the row index
    {
      const int out8 = (d.widths[1] + 7) & ~7;
      const float* W = s_f;
      const float* Bv = W + n * out8;
      if (ty * 8 < out8) {
        float base[8];
        {
          const float4 b0 = *reinterpret_cast<const float4*>(Bv + ty * 8), b1 = *reinterpret_cast<const float4*>(Bv + ty * 8 + 4);
          base[0] = b0.x; base[1] = b0.y; base[2] = b0.z; base[3] = b0.w;
          base[4] = b1.x; base[5] = b1.y; base[6] = b1.z; base[7] = b1.w;
        }
        const uint64_t rowbase = row0 + (uint64_t)tx * 4;  // bits 0, 1 and 6 are clear
        for (int p = 2; p < n; ++p) {
          if (p == 6 || !((rowbase >> p) & 1ull)) continue;
          const float* wr = W + (n - 1 - p) * out8 + ty * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
          base[0] += w0.x; base[1] += w0.y; base[2] += w0.z; base[3] += w0.w;
          base[4] += w1.x; base[5] += w1.y; base[6] += w1.z; base[7] += w1.w;
        }
        float wp[3][8];  // weight rows of index bits 0, 1, 6 (zero when the model has fewer bits)
        const int pb[3] = {0, 1, 6};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          if (pb[q] < n) {
            const float* wr = W + (n - 1 - pb[q]) * out8 + ty * 8;
            const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
            wp[q][0] = w0.x; wp[q][1] = w0.y; wp[q][2] = w0.z; wp[q][3] = w0.w;
            wp[q][4] = w1.x; wp[q][5] = w1.y; wp[q][6] = w1.z; wp[q][7] = w1.w;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) wp[q][j] = 0.f;
          }
        }
        const int act = d.act[0];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            float y = base[j];
            if (r & 1) y += wp[0][j];
            if (r & 2) y += wp[1][j];
            if (r & 4) y += wp[2][j];
            v[r] = mlp_act(y, act);
          }
          float* o = buf + (ty * 8 + j) * kMlpRows + tx * 4;
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(o + 64) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    }
    __syncthreads();
    int off = n * ((d.widths[1] + 7) & ~7) + ((d.widths[1] + 7) & ~7);
    // ---- middle layers: [128 x in] x [in x out], 8 x 8 micro-tiles, in place through registers
    for (int l = 1; l + 1 < d.n_layers; ++l) {
      const int in = d.widths[l], out = d.widths[l + 1];
      const int out8 = (out + 7) & ~7;
      const float* W = s_f + off;
      const float* Bv = W + in * out8;
      const bool mine = ty * 8 < out8;
      // Packed accumulators: acc2[r][jj] = outputs (2jj, 2jj+1) of row r.  One FFMA2 (sm_100 packed fp32: two
      // independent round-to-nearest FMAs, the activation broadcast as an operand modifier) per pair, so a
      // k-step is 32 issue slots for 64 FMAs and the loads / address updates issue in the FMA pipe's shadow.
      float2 acc2[8][4];
      if (mine) {
        const float4 b0 = *reinterpret_cast<const float4*>(Bv + ty * 8), b1 = *reinterpret_cast<const float4*>(Bv + ty * 8 + 4);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          acc2[r][0] = make_float2(b0.x, b0.y); acc2[r][1] = make_float2(b0.z, b0.w);
          acc2[r][2] = make_float2(b1.x, b1.y); acc2[r][3] = make_float2(b1.z, b1.w);
        }
#pragma unroll 2
        for (int k = 0; k < in; ++k) {
          const float4 x0 = *reinterpret_cast<const float4*>(buf + k * kMlpRows + tx * 4);
          const float4 x1 = *reinterpret_cast<const float4*>(buf + k * kMlpRows + tx * 4 + 64);
          const float4 w0 = *reinterpret_cast<const float4*>(W + k * out8 + ty * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(W + k * out8 + ty * 8 + 4);
          const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
          const float2 ws2[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y),
                                 make_float2(w1.z, w1.w)};
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc2[r][jj] = __ffma2_rn(make_float2(xs[r], xs[r]), ws2[jj], acc2[r][jj]);
        }
      }
      __syncthreads();  // every thread has read its inputs: the buffer can take the outputs
      if (mine) {
        const int act = d.act[l];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float* o = buf + (ty * 8 + j) * kMlpRows + tx * 4;
          auto A = [&](int r) { return (j & 1) ? acc2[r][j >> 1].y : acc2[r][j >> 1].x; };
          *reinterpret_cast<float4*>(o) = make_float4(mlp_act(A(0), act), mlp_act(A(1), act), mlp_act(A(2), act),
                                                      mlp_act(A(3), act));
          *reinterpret_cast<float4*>(o + 64) = make_float4(mlp_act(A(4), act), mlp_act(A(5), act), mlp_act(A(6), act),
                                                           mlp_act(A(7), act));
        }
      }
      off += in * out8 + out8;
      __syncthreads();
    }
    // ---- last layer: one output per row, one row per thread
    {
      const int l = d.n_layers - 1;
      const int in = d.widths[l];
      const int out8 = mlp_out_pad(d, l);
      const float* W = s_f + off;
      const uint64_t row = row0 + (uint64_t)tid;
      float e0 = W[in * out8], e1 = 0.f, e2 = 0.f, e3 = 0.f;
      int k = 0;
      for (; k + 3 < in; k += 4) {
        e0 = fmaf(buf[k * kMlpRows + tid], W[k * out8], e0);
        e1 = fmaf(buf[(k + 1) * kMlpRows + tid], W[(k + 1) * out8], e1);
        e2 = fmaf(buf[(k + 2) * kMlpRows + tid], W[(k + 2) * out8], e2);
        e3 = fmaf(buf[(k + 3) * kMlpRows + tid], W[(k + 3) * out8], e3);
      }
      for (; k < in; ++k) e0 = fmaf(buf[k * kMlpRows + tid], W[k * out8], e0);
      if (row >= lo && row < hi) {
        const float lg = -mlp_act((e0 + e1) + (e2 + e3), d.act[l]);
        if (logits) logits[row - lo] = lg;
        stat_add(acc_st, (double)lg);
      }
    }
  }
  acc_st = stat_warp(acc_st);
  if ((tid & 31) == 0) s_st[tid >> 5] = acc_st;
  __syncthreads();
  if (tid < 32) {
    Stat v;
    v.m = 0.0; v.s = 0.0; v.t = 0.0;
    if (tid < kMlpThreads / 32) v = s_st[tid];
    v = stat_warp(v);
    if (tid == 0) partial[blockIdx.x] = v;
  }
}

// ------------------------------------------------------------------ parity-energy sweep, Walsh-Hadamard tiles
// E(i) = sum_t theta_t (-1)^{parity(i & mask_t)}.  Inside a tile of 256 consecutive rows only the low
// 8 index bits vary, so E restricted to the tile is the 256-point Walsh-Hadamard transform of the
// sparse vector c[m] = sum over the terms whose low mask byte is m of theta_t (-1)^{parity(hi & mask_t)}.
// Per tile: every term is touched once (not once per row) and the transform costs 8 butterflies per
// row.  Bucket sums run in term order, so results do not depend on scheduling (seeded sampling from
// the logits stays exactly repeatable).
constexpr int kParThreads = 256;

__global__ void __launch_bounds__(kParThreads) ebm_parity_sweep_kernel(const __grid_constant__ EnergyArgs ea, uint64_t lo,
                                                                       uint64_t hi, float* __restrict__ logits,
                                                                       Stat* __restrict__ partial) {
  extern __shared__ __align__(16) float s_f[];
  __shared__ Stat s_st[kParThreads / 32];
  __shared__ int s_start[kParThreads + 1];
  __shared__ float s_x[kParThreads];
  const qhbm_energy_desc_t& d = ea.d;
  const int nt = d.n_terms;
  const int tid = threadIdx.x;
  float* s_theta = s_f;
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_f + nt);
  int* s_order = reinterpret_cast<int*>(s_f + 2 * nt);
  for (int i = tid; i < nt; i += kParThreads) {
    s_theta[i] = d.d_theta[i];
    s_mask[i] = d.d_masks[i];
  }
  __syncthreads();
  // terms grouped by the low byte of their mask, in term order inside a group (stable, deterministic)
  {
    int cnt = 0;
    for (int t = 0; t < nt; ++t) cnt += (int)(s_mask[t] & 255u) == tid;
    s_x[tid] = __int_as_float(cnt);
    __syncthreads();
    if (tid == 0) {
      int run = 0;
      for (int b = 0; b < kParThreads; ++b) {
        s_start[b] = run;
        run += __float_as_int(s_x[b]);
      }
      s_start[kParThreads] = run;
    }
    __syncthreads();
    int pos = s_start[tid];
    for (int t = 0; t < nt; ++t)
      if ((int)(s_mask[t] & 255u) == tid) s_order[pos++] = t;
    __syncthreads();
  }
  const int k0 = s_start[tid], k1 = s_start[tid + 1];
  const int n_b0 = s_start[1];  // bucket 0 occupies order[0 .. n_b0)
  const uint64_t t_first = lo >> 8, t_last = (hi - 1) >> 8;
  Stat acc_st;
  acc_st.m = 0.0; acc_st.s = 0.0; acc_st.t = 0.0;
  for (uint64_t tile = t_first + blockIdx.x; tile <= t_last; tile += gridDim.x) {
    const uint32_t hbits = (uint32_t)(tile << 8);  // parity energies have at most 32 bits
    float c = 0.f;
    if (tid != 0) {
      for (int k = k0; k < k1; ++k) {
        const int t = s_order[k];
        const uint32_t sgn = (uint32_t)(__popc(hbits & s_mask[t]) & 1) << 31;
        c += __uint_as_float(__float_as_uint(s_theta[t]) ^ sgn);
      }
    }
    // bucket 0 (terms that do not touch the low byte: most of them) is summed by the whole CTA in a
    // fixed order: strided partials, shuffle tree, then the eight warp sums
    {
      float p0 = 0.f;
      for (int k = tid; k < n_b0; k += kParThreads) {
        const int t = s_order[k];
        const uint32_t sgn = (uint32_t)(__popc(hbits & s_mask[t]) & 1) << 31;
        p0 += __uint_as_float(__float_as_uint(s_theta[t]) ^ sgn);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) p0 += __shfl_xor_sync(0xffffffffu, p0, o);
      __syncthreads();  // previous tile's last exchange is read
      if ((tid & 31) == 0) s_x[tid >> 5] = p0;
      __syncthreads();
      if (tid == 0) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < kParThreads / 32; ++w) sum += s_x[w];
        c = sum;
      }
    }
    // 256-point Walsh-Hadamard transform: bits 0-4 by shuffles, bits 5-7 through shared memory
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      const float o = __shfl_xor_sync(0xffffffffu, c, 1 << b);
      c = (tid & (1 << b)) ? o - c : c + o;
    }
#pragma unroll
    for (int b = 5; b < 8; ++b) {
      __syncthreads();
      s_x[tid] = c;
      __syncthreads();
      const float o = s_x[tid ^ (1 << b)];
      c = (tid & (1 << b)) ? o - c : c + o;
    }
    const uint64_t row = (tile << 8) | (uint64_t)tid;
    if (row >= lo && row < hi) {
      const float lg = -c;
      if (logits) logits[row - lo] = lg;
      stat_add(acc_st, (double)lg);
    }
  }
  acc_st = stat_warp(acc_st);
  if ((tid & 31) == 0) s_st[tid >> 5] = acc_st;
  __syncthreads();
  if (tid < 32) {
    Stat v;
    v.m = 0.0; v.s = 0.0; v.t = 0.0;
    if (tid < kParThreads / 32) v = s_st[tid];
    v = stat_warp(v);
    if (tid == 0) partial[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) stat_final_kernel(const Stat* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ Stat s_st[8];
  Stat acc;
  acc.m = 0.0; acc.s = 0.0; acc.t = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc = stat_merge(acc, partial[i]);
  acc = stat_warp(acc);
  if ((threadIdx.x & 31) == 0) s_st[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    Stat v;
    v.m = 0.0; v.s = 0.0; v.t = 0.0;
    if (threadIdx.x < 8) v = s_st[threadIdx.x];
    v = stat_warp(v);
    if (threadIdx.x == 0) { out[0] = v.m; out[1] = v.s; out[2] = v.t; }
  }
}

// ------------------------------------------------------------------ categorical sampling
constexpr int kCatBlock = 256;  // rows per prefix block
__global__ void __launch_bounds__(256) cat_max_kernel(const float* __restrict__ logits, int64_t n, float* __restrict__ gmax) {
  float m = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, logits[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (m == 0.f) m = 0.f;  // -0.0 would order below every negative float in the int trick
  if ((threadIdx.x & 31) == 0) {
    // float atomic max via int ordering trick
    int* a = reinterpret_cast<int*>(gmax);
    if (m >= 0.f) atomicMax(a, __float_as_int(m));
    else atomicMin(reinterpret_cast<unsigned int*>(a), __float_as_uint(m));
  }
}
// Block sums of exp(l - max) over 256 rows, plus the 8 sub-block sums over 32 rows each (`sub`): a sample
// then scans 8 sub-block sums and at most 32 rows instead of up to 256 rows.  The block sum is the
// sequential sum of its sub-block sums, the order the sample kernel repeats.
__global__ void __launch_bounds__(kCatBlock) cat_blocksum_kernel(const float* __restrict__ logits, int64_t n,
                                                                 const float* __restrict__ gmax, double* __restrict__ bsum,
                                                                 double* __restrict__ sub) {
  __shared__ double s_w[kCatBlock / 32];
  const int64_t i = (int64_t)blockIdx.x * kCatBlock + threadIdx.x;
  double v = i < n ? exp((double)logits[i] - (double)*gmax) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) {
    s_w[threadIdx.x >> 5] = v;
    sub[(int64_t)blockIdx.x * (kCatBlock / 32) + (threadIdx.x >> 5)] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < kCatBlock / 32; ++k) s += s_w[k];
    bsum[blockIdx.x] = s;
  }
}
// exclusive scan of block sums (float64) by one block; total -> bsum[nblocks]
__global__ void __launch_bounds__(1024) cat_scan_kernel(double* __restrict__ bsum, int64_t nblocks) {
  __shared__ double s_w[32];
  __shared__ double s_carry;
  if (threadIdx.x == 0) s_carry = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t b0 = 0; b0 < nblocks; b0 += 4096) {  // four consecutive entries per thread
    const int64_t i4 = b0 + 4 * (int64_t)threadIdx.x;
    double f4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) f4[k] = i4 + k < nblocks ? bsum[i4 + k] : 0.0;
    const double f = ((f4[0] + f4[1]) + f4[2]) + f4[3];
    double v = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) s_w[wid] = v;
    __syncthreads();
    if (wid == 0) {
      double x = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
      }
      s_w[lane] = x;
    }
    __syncthreads();
    const double incl = v + (wid ? s_w[wid - 1] : 0.0) + s_carry;
    double run = incl - f;  // exclusive prefix of this thread's first entry
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i4 + k < nblocks) bsum[i4 + k] = run;
      run += f4[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[nblocks] = s_carry;
}
__global__ void cat_sample_kernel(const float* __restrict__ logits, int64_t n, const float* __restrict__ gmax,
                                  const double* __restrict__ bpre, const double* __restrict__ sub, int64_t nblocks,
                                  uint64_t row_offset,
                                  uint64_t seed0, uint64_t seed1, uint64_t first, int64_t n_samples,
                                  uint64_t* __restrict__ out, double mass_begin, double mass_end, double mass_total) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_samples) return;
  const Philox ph = make_philox(seed0, seed1);
  const uint4 r = ph(first + (uint64_t)k, 0x43415453u);
  double target;
  if (mass_total > 0.0) {
    // one shard of a row range split over ranks: sample k belongs to the rank whose mass interval
    // [mass_begin, mass_end) holds its point of the GLOBAL cumulative mass; the others leave out[k] alone
    const double t = u01_53(r.x, r.y) * mass_total;
    if (!(t >= mass_begin && t < mass_end)) return;
    target = t - mass_begin;
  } else {
    target = u01_53(r.x, r.y) * bpre[nblocks];
  }
  // largest block b with bpre[b] <= target
  int64_t lo = 0, hi = nblocks - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (bpre[mid] <= target) lo = mid; else hi = mid - 1;
  }
  const double m = (double)*gmax;
  double acc = bpre[lo];
  // the 32-row sub-block inside block lo (the last one if rounding pushed the target past the block)
  int w = 0;
  for (; w < kCatBlock / 32 - 1; ++w) {
    const double sw = sub[lo * (kCatBlock / 32) + w];
    if (target < acc + sw) break;
    acc += sw;
  }
  const int64_t i0 = min(n - 1, lo * kCatBlock + 32 * (int64_t)w), i1 = min(n, lo * kCatBlock + 32 * (int64_t)(w + 1));
  int64_t pick = i1 - 1;
  int64_t last_pos = -1;
  for (int64_t i = i0; i < i1; ++i) {
    const double wt = exp((double)logits[i] - m);
    if (wt > 0.0) last_pos = i;
    acc += wt;
    if (target < acc) { pick = i; last_pos = -2; break; }
  }
  if (last_pos >= 0) pick = last_pos;  // rounding pushed the target past the sub-block: last positive weight
  out[k] = row_offset + (uint64_t)pick;
}

__global__ void bern_sample_kernel(const float* __restrict__ logits, int n_bits, ShiftTable st, uint64_t seed0,
                                   uint64_t seed1, uint64_t first, int64_t n_samples, uint64_t* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_samples) return;
  const Philox ph = make_philox(seed0, seed1);
  uint64_t key = 0;
  for (int j0 = 0; j0 < n_bits; j0 += 4) {
    const uint4 r = ph(first + (uint64_t)k, 0x42524E00u + (uint32_t)(j0 >> 2));
    const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + q;
      if (j < n_bits) {
        const float p1 = 1.f / (1.f + expf(-logits[j]));
        if (u01_24(rv[q]) < p1) key |= 1ull << st.s[j];
      }
    }
  }
  out[k] = key;
}

static ShiftTable make_shift(const int32_t* h_shift, int n_bits) {
  if (n_bits < 1 || n_bits > 63) throw std::runtime_error("n_bits must be in [1, 63]");
  if (!h_shift) throw std::runtime_error("h_shift is null");
  ShiftTable st;
  for (int j = 0; j < 64; ++j) st.s[j] = 0;
  for (int j = 0; j < n_bits; ++j) {
    if (h_shift[j] < 0 || h_shift[j] > 62) throw std::runtime_error("shift out of range");
    st.s[j] = (int8_t)h_shift[j];
  }
  return st;
}

static uint64_t next_pow2(uint64_t x) {
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}
static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static UniqueWs carve_unique(void* ws, int64_t n) {
  UniqueWs w;
  w.cap = std::max<uint64_t>(1024, next_pow2((uint64_t)n * 2));
  const int64_t nblocks = (n + kScanBlock - 1) / kScanBlock;
  char* p = reinterpret_cast<char*>(ws);
  w.tkeys = reinterpret_cast<unsigned long long*>(p); p += align_up(w.cap * 8);
  w.tfirst = reinterpret_cast<int32_t*>(p); p += align_up((w.cap + 1) * 4);
  w.tcount = reinterpret_cast<int32_t*>(p); p += align_up((w.cap + 1) * 4);
  w.slot = reinterpret_cast<int32_t*>(p); p += align_up((size_t)n * 4);
  w.rank = reinterpret_cast<int32_t*>(p); p += align_up((size_t)n * 4);
  w.bsum = reinterpret_cast<int32_t*>(p); p += align_up((size_t)(nblocks + 1) * 4);
  return w;
}

}  // namespace qhbm

using namespace qhbm;

namespace {
// 28 KB of scratch per (device, stream): work on one stream is ordered, so the buffer can be reused by
// the next sweep on that stream; different streams and devices never share one.
void* sweep_scratch(cudaStream_t s) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, void*> cache;
  int dev = 0;
  QHBM_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  void*& p = cache[{dev, s}];
  if (!p) QHBM_CUDA(cudaMalloc(&p, sizeof(Stat) * 148 * 8));
  return p;
}
}  // namespace

extern "C" {

int qhbm_pack_bits(const int8_t* d_bits, int64_t n_rows, int32_t n_bits, const int32_t* h_shift, uint64_t* d_keys,
                   void* stream) {
  return guarded([&] {
    const ShiftTable st = make_shift(h_shift, n_bits);
    if (n_rows <= 0) return;
    pack_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_bits, n_rows, n_bits, st, d_keys);
    QHBM_CUDA(cudaGetLastError());
  });
}
int qhbm_unpack_bits(const uint64_t* d_keys, int64_t n_rows, int32_t n_bits, const int32_t* h_shift, int8_t* d_bits,
                     void* stream) {
  return guarded([&] {
    const ShiftTable st = make_shift(h_shift, n_bits);
    if (n_rows <= 0) return;
    const int64_t tot = n_rows * n_bits;
    unpack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_keys, n_rows, n_bits, st, d_bits);
    QHBM_CUDA(cudaGetLastError());
  });
}

int64_t qhbm_unique_workspace_bytes(int64_t n_rows) {
  if (n_rows < 0) return -1;
  const uint64_t cap = std::max<uint64_t>(1024, next_pow2((uint64_t)n_rows * 2));
  const int64_t nblocks = (n_rows + kScanBlock - 1) / kScanBlock;
  return (int64_t)(align_up(cap * 8) + 2 * align_up((cap + 1) * 4) + 2 * align_up((size_t)n_rows * 4) +
                   align_up((size_t)(nblocks + 1) * 4) + 256);
}

int qhbm_unique_with_counts(const uint64_t* d_keys, int64_t n_rows, uint64_t* d_unique, int32_t* d_idx,
                            int32_t* d_count, int64_t* d_n_unique, void* d_workspace, void* stream) {
  return guarded([&] {
    cudaStream_t s = (cudaStream_t)stream;
    if (n_rows < 0 || n_rows > 0x7fffffff) throw std::runtime_error("n_rows out of range");
    if (n_rows == 0) {
      QHBM_CUDA(cudaMemsetAsync(d_n_unique, 0, sizeof(int64_t), s));
      return;
    }
    if (!d_workspace) throw std::runtime_error("workspace is null");
    UniqueWs w = carve_unique(d_workspace, n_rows);
    const int64_t nblocks = (n_rows + kScanBlock - 1) / kScanBlock;
    const unsigned g = (unsigned)((n_rows + 255) / 256);
    uq_init<<<(unsigned)((w.cap + 256) / 256), 256, 0, s>>>(w);
    uq_insert<<<g, 256, 0, s>>>(d_keys, n_rows, w);
    uq_flag_scan<<<(unsigned)nblocks, kScanBlock, 0, s>>>(n_rows, w);
    uq_scan_blocks<<<1, 1024, 0, s>>>(nblocks, w, d_n_unique);
    uq_emit<<<g, 256, 0, s>>>(d_keys, n_rows, w, d_unique, d_count);
    uq_index<<<g, 256, 0, s>>>(n_rows, w, d_idx);
    QHBM_CUDA(cudaGetLastError());
  });
}

int qhbm_segment_sum(const float* d_vals, const int32_t* d_idx, int64_t n_rows, int32_t width, float* d_out,
                     int64_t n_unique, void* stream) {
  return guarded([&] {
    cudaStream_t s = (cudaStream_t)stream;
    if (width < 1) throw std::runtime_error("width must be >= 1");
    QHBM_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float) * n_unique * width, s));
    const int64_t tot = n_rows * width;
    if (tot <= 0) return;
    segsum_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_vals, d_idx, n_rows, width, d_out);
    QHBM_CUDA(cudaGetLastError());
  });
}

int qhbm_weighted_sum(const int32_t* d_counts, const float* d_vals, int64_t n_rows, int32_t width, double* d_out,
                      void* stream) {
  return guarded([&] {
    cudaStream_t s = (cudaStream_t)stream;
    if (width < 0) throw std::runtime_error("width must be >= 0");
    QHBM_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * (width + 1), s));
    if (n_rows <= 0) return;
    const int threads = std::min(1024, ((width + 1 + 31) / 32) * 32);
    wsum_kernel<<<(unsigned)((n_rows + kWsumStrip - 1) / kWsumStrip), threads, 0, s>>>(d_counts, d_vals, n_rows, width, d_out);
    QHBM_CUDA(cudaGetLastError());
  });
}

int qhbm_score_gradient(const uint64_t* d_keys, const int32_t* d_counts, int64_t n_rows, const float* d_vals,
                        int32_t width, const float* d_upstream, const float* d_average, const int32_t* d_masks,
                        int32_t n_terms, const double* d_total_count, float scale, float* d_grad_theta,
                        float* d_workspace, void* stream) {
  return guarded([&] {
    if (n_rows < 0 || width < 1 || n_terms < 0) throw std::runtime_error("bad sizes");
    if (!d_workspace) throw std::runtime_error("workspace is null");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_terms == 0) return;
    if (n_rows == 0) {
      QHBM_CUDA(cudaMemsetAsync(d_grad_theta, 0, sizeof(float) * n_terms, s));
      return;
    }
    float* s0 = d_workspace;
    float* c = d_workspace + 4;
    score_c_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, s>>>(d_vals, n_rows, width, d_upstream, d_average, c, s0);
    score_term_kernel<<<(unsigned)n_terms, 256, 0, s>>>(d_keys, d_counts, n_rows, c, s0, d_masks, d_total_count, scale,
                                                        d_grad_theta);
    QHBM_CUDA(cudaGetLastError());
  });
}

int qhbm_energy_rows(const qhbm_energy_desc_t* e, const uint64_t* d_keys, int64_t n_rows, float* d_energy,
                     void* stream) {
  return guarded([&] {
    validate_energy(e);
    if (n_rows <= 0) return;
    EnergyArgs ea;
    ea.d = *e;
    const size_t smem = energy_smem_bytes(*e);
    QHBM_CUDA(cudaFuncSetAttribute(energy_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int64_t blocks = std::min<int64_t>((n_rows + kEnergyThreads - 1) / kEnergyThreads, 148 * 8);
    energy_rows_kernel<<<(unsigned)blocks, kEnergyThreads, smem, (cudaStream_t)stream>>>(ea, d_keys, n_rows, d_energy);
    QHBM_CUDA(cudaGetLastError());
  });
}


int qhbm_ebm_sweep(const qhbm_energy_desc_t* e, uint64_t lo, uint64_t hi, float* d_logits, double* d_stats,
                   void* stream) {
  return guarded([&] {
    validate_energy(e);
    if (hi < lo) throw std::runtime_error("hi < lo");
    if (e->n_bits < 63 && hi > (1ull << e->n_bits)) throw std::runtime_error("row range exceeds 2^n_bits");
    if (!d_stats) throw std::runtime_error("d_stats is null");
    cudaStream_t s = (cudaStream_t)stream;
    EnergyArgs ea;
    ea.d = *e;
    const size_t smem = energy_smem_bytes(*e);
    const uint64_t rows = hi - lo;
    int blocks = (int)std::min<uint64_t>((rows + kEnergyThreads - 1) / kEnergyThreads, 148 * 8);
    blocks = std::max(blocks, 1);
    void* partial = sweep_scratch(s);  // per-CTA partial statistics
    // The kernel is chosen from the PROBLEM size (2^n_bits rows), never from this call's row count: a
    // rank that sweeps one shard of the range must produce bit-identical logits to a single sweep of the
    // whole range, or the sharded sampler would not reproduce the single-GPU draw.
    const bool big = e->n_bits >= 12 && rows >= 1;
    if (e->kind == QHBM_ENERGY_MLP && e->n_layers >= 2 && big) {
      // register-tiled dense-stack kernel: weights + one [64][128] activation buffer in shared memory
      const size_t msmem = ((mlp_sweep_weight_floats(*e) * sizeof(float) + 15) & ~(size_t)15) +
                           sizeof(float) * kMaxWidth * kMlpRows;
      QHBM_CUDA(cudaFuncSetAttribute(ebm_mlp_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
      const uint64_t ntiles = (hi - 1) / kMlpRows - lo / kMlpRows + 1;
      blocks = (int)std::min<uint64_t>(ntiles, 148 * 4);  // four 57 KB CTAs per SM
      ebm_mlp_sweep_kernel<<<blocks, kMlpThreads, msmem, s>>>(ea, lo, hi, d_logits, (Stat*)partial);
    } else if (e->kind != QHBM_ENERGY_MLP && big && (size_t)e->n_terms * 12 <= 200 * 1024) {
      // Walsh-Hadamard tiles of 256 rows: theta, masks and the bucket order in shared memory
      const size_t psmem = (size_t)std::max(e->n_terms, 1) * 12;
      QHBM_CUDA(cudaFuncSetAttribute(ebm_parity_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
      const uint64_t ntiles = ((hi - 1) >> 8) - (lo >> 8) + 1;
      blocks = (int)std::min<uint64_t>(ntiles, 148 * 8);
      ebm_parity_sweep_kernel<<<blocks, kParThreads, psmem, s>>>(ea, lo, hi, d_logits, (Stat*)partial);
    } else {
      QHBM_CUDA(cudaFuncSetAttribute(ebm_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      ebm_sweep_kernel<<<blocks, kEnergyThreads, smem, s>>>(ea, lo, hi, d_logits, (Stat*)partial);
    }
    stat_final_kernel<<<1, 256, 0, s>>>((const Stat*)partial, blocks, d_stats);
    QHBM_CUDA(cudaGetLastError());
  });
}

int64_t qhbm_sample_workspace_bytes(int64_t n_rows) {
  if (n_rows < 0) return -1;
  const int64_t nblocks = (n_rows + kCatBlock - 1) / kCatBlock;
  return 256 + 8 * (nblocks + 2) + 8 * (kCatBlock / 32) * nblocks;  // max | block prefix + total | sub-block sums
}

static void cat_prepare(const float* d_logits, int64_t n_rows, bool use_given_max, float given_max, void* d_workspace,
                        cudaStream_t s) {
  if (n_rows < 1) throw std::runtime_error("n_rows must be >= 1");
  if (!d_workspace) throw std::runtime_error("workspace is null");
  float* gmax = reinterpret_cast<float*>(d_workspace);
  double* bsum = reinterpret_cast<double*>(reinterpret_cast<char*>(d_workspace) + 256);
  const int64_t nblocks = (n_rows + kCatBlock - 1) / kCatBlock;
  const float init = use_given_max ? given_max : -INFINITY;
  QHBM_CUDA(cudaMemcpyAsync(gmax, &init, sizeof(float), cudaMemcpyHostToDevice, s));
  if (!use_given_max)
    cat_max_kernel<<<(unsigned)std::min<int64_t>((n_rows + 255) / 256, 148 * 8), 256, 0, s>>>(d_logits, n_rows, gmax);
  cat_blocksum_kernel<<<(unsigned)nblocks, kCatBlock, 0, s>>>(d_logits, n_rows, gmax, bsum, bsum + nblocks + 2);
  cat_scan_kernel<<<1, 1024, 0, s>>>(bsum, nblocks);
  QHBM_CUDA(cudaGetLastError());
}

static void cat_draw(const float* d_logits, int64_t n_rows, uint64_t row_offset, const void* d_workspace,
                     double mass_begin, double mass_end, double mass_total, uint64_t seed0, uint64_t seed1,
                     uint64_t first_sample, int64_t n_samples, uint64_t* d_samples, cudaStream_t s) {
  if (n_samples < 0) throw std::runtime_error("n_samples must be >= 0");
  if (n_samples == 0) return;
  const float* gmax = reinterpret_cast<const float*>(d_workspace);
  const double* bsum = reinterpret_cast<const double*>(reinterpret_cast<const char*>(d_workspace) + 256);
  const int64_t nblocks = (n_rows + kCatBlock - 1) / kCatBlock;
  cat_sample_kernel<<<(unsigned)((n_samples + 127) / 128), 128, 0, s>>>(d_logits, n_rows, gmax, bsum, bsum + nblocks + 2, nblocks, row_offset,
                                                                       seed0, seed1, first_sample, n_samples, d_samples,
                                                                       mass_begin, mass_end, mass_total);
  QHBM_CUDA(cudaGetLastError());
}

int qhbm_categorical_sample(const float* d_logits, int64_t n_rows, uint64_t row_offset, uint64_t seed0,
                            uint64_t seed1, uint64_t first_sample, int64_t n_samples, uint64_t* d_samples,
                            void* d_workspace, void* stream) {
  return guarded([&] {
    cat_prepare(d_logits, n_rows, false, 0.f, d_workspace, (cudaStream_t)stream);
    cat_draw(d_logits, n_rows, row_offset, d_workspace, 0.0, 0.0, 0.0, seed0, seed1, first_sample, n_samples, d_samples,
             (cudaStream_t)stream);
  });
}

int qhbm_categorical_prepare(const float* d_logits, int64_t n_rows, int32_t use_given_max, float given_max,
                             void* d_workspace, void* stream) {
  return guarded([&] { cat_prepare(d_logits, n_rows, use_given_max != 0, given_max, d_workspace, (cudaStream_t)stream); });
}

int qhbm_categorical_draw(const float* d_logits, int64_t n_rows, uint64_t row_offset, const void* d_workspace,
                          double mass_begin, double mass_end, double mass_total, uint64_t seed0, uint64_t seed1,
                          uint64_t first_sample, int64_t n_samples, uint64_t* d_samples, void* stream) {
  return guarded([&] {
    if (n_rows < 1) throw std::runtime_error("n_rows must be >= 1");
    if (!d_workspace) throw std::runtime_error("workspace is null");
    cat_draw(d_logits, n_rows, row_offset, d_workspace, mass_begin, mass_end, mass_total, seed0, seed1, first_sample,
             n_samples, d_samples, (cudaStream_t)stream);
  });
}

int qhbm_bernoulli_sample(const float* d_logits, int32_t n_bits, const int32_t* h_shift, uint64_t seed0,
                          uint64_t seed1, uint64_t first_sample, int64_t n_samples, uint64_t* d_samples,
                          void* stream) {
  return guarded([&] {
    const ShiftTable st = make_shift(h_shift, n_bits);
    if (n_samples <= 0) return;
    bern_sample_kernel<<<(unsigned)((n_samples + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_logits, n_bits, st, seed0, seed1, first_sample, n_samples, d_samples);
    QHBM_CUDA(cudaGetLastError());
  });
}

}  // extern "C"
