// philox.h -- counter-based random numbers shared by the sampling kernels (ebm.cu, measure.cu).
#pragma once
#include <stdint.h>

namespace qhbm {

// ------------------------------------------------------------------ Philox4x32-10
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ uint4 operator()(uint64_t counter, uint32_t stream_hi) const {
    uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32), c2 = stream_hi, c3 = 0x9E3779B9u;
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
__device__ __forceinline__ Philox make_philox(uint64_t seed0, uint64_t seed1) {
  Philox p;
  p.k0 = (uint32_t)seed0 ^ (uint32_t)(seed1 >> 32);
  p.k1 = (uint32_t)(seed0 >> 32) ^ (uint32_t)seed1 * 0x85EBCA6Bu;
  return p;
}
__device__ __forceinline__ double u01_53(uint32_t hi, uint32_t lo) {
  const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ float u01_24(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

}  // namespace qhbm
