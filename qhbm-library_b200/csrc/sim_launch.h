// sim_launch.h -- launch entry points of the sweep-kernel instantiations.  The general (GEN = true) and
// lean (GEN = false) instantiations live in two translation units (sim.cu, sim_lean.cu) so that they
// compile in parallel; both expose the same two functions.
#pragma once
#include <cuda_runtime.h>

#include "sim_kernels.cuh"

namespace qhbm {

// K in {4, 5}; dense = the three-CTAs-per-SM forward variant (K = 4, forward only).
void launch_sweep_gen(int K, bool adj, bool dense, const KernelArgs& ka, unsigned grid, int threads, size_t smem,
                      cudaStream_t s);
void launch_sweep_lean(int K, bool adj, bool dense, const KernelArgs& ka, unsigned grid, int threads, size_t smem,
                       cudaStream_t s);
// Opt in to large dynamic shared memory for every instantiation of the unit (sticky per device).
void allow_large_smem_gen();
void allow_large_smem_lean();

// Shared body of the two units.
template <bool GEN>
inline void launch_sweep_impl(int K, bool adj, bool dense, const KernelArgs& ka, unsigned grid, int threads,
                              size_t smem, cudaStream_t s) {
  if (dense) sweep_kernel<4, false, true, GEN><<<grid, threads, smem, s>>>(ka);
  else if (K == 4 && adj) sweep_kernel<4, true, false, GEN><<<grid, threads, smem, s>>>(ka);
  else if (K == 4) sweep_kernel<4, false, false, GEN><<<grid, threads, smem, s>>>(ka);
  else if (adj) sweep_kernel<5, true, false, GEN><<<grid, threads, smem, s>>>(ka);
  else sweep_kernel<5, false, false, GEN><<<grid, threads, smem, s>>>(ka);
}
template <bool GEN>
inline cudaError_t allow_large_smem_impl() {
  int dev = 0, optin = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  const void* fns[5] = {reinterpret_cast<const void*>(sweep_kernel<4, true, false, GEN>),
                        reinterpret_cast<const void*>(sweep_kernel<4, false, false, GEN>),
                        reinterpret_cast<const void*>(sweep_kernel<5, true, false, GEN>),
                        reinterpret_cast<const void*>(sweep_kernel<5, false, false, GEN>),
                        reinterpret_cast<const void*>(sweep_kernel<4, false, true, GEN>)};
  for (const void* fn : fns) {
    cudaFuncAttributes attr;
    e = cudaFuncGetAttributes(&attr, fn);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)attr.sharedSizeBytes);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace qhbm
