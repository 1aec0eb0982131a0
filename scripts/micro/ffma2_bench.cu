// Microbenchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, and of the
// mixed sequences the sweep kernel would use.  Build: nvcc -arch=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void k(float* out, int iters, float c, float s) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 cc = make_float2(c, c), ss = make_float2(s, -s);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // 2 scalar FFMA per element
        a[i].x = fmaf(a[i].x, c, s);
        a[i].y = fmaf(a[i].y, c, -s);
      } else if (MODE == 1) {  // 1 FFMA2 per element
        a[i] = __ffma2_rn(a[i], cc, ss);
      } else if (MODE == 2) {  // FFMA2 + FMUL2 alternating
        a[i] = __ffma2_rn(a[i], cc, ss);
        a[i] = __fmul2_rn(a[i], cc);
      } else {  // scalar equivalent of MODE 2
        a[i].x = fmaf(a[i].x, c, s);
        a[i].y = fmaf(a[i].y, c, -s);
        a[i].x *= c;
        a[i].y *= c;
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, int threads, int blocks_per_sm, float* d) {
  int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int sms = 148;
  k<MODE><<<sms * blocks_per_sm, threads>>>(d, 16, 0.999f, 0.001f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<sms * blocks_per_sm, threads>>>(d, iters, 0.999f, 0.001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double elems = (double)sms * blocks_per_sm * threads * 8.0 * iters;  // float2 elements updated
  const double fl = elems * 2 * 2 * (MODE >= 2 ? 1.5 : 1.0);
  printf("%-28s threads/SM %4d: %7.3f ms  %7.2f Gelem/s  %6.2f TFLOP/s\n", name, threads * blocks_per_sm, ms,
         elems / ms * 1e-6, fl / ms * 1e-9);
}

int main() {
  float* d;
  cudaMalloc(&d, 148 * 16 * 1024 * sizeof(float));
  for (int bps : {2, 4, 8}) {
    run<0>("scalar FFMA x2", 256, bps, d);
    run<1>("FFMA2", 256, bps, d);
    run<3>("scalar FFMA x2 + FMUL x2", 256, bps, d);
    run<2>("FFMA2 + FMUL2", 256, bps, d);
  }
  return 0;
}
