// sim_lean.cu -- the GEN = false instantiations of the sweep kernels (plans without general-matrix ops:
// X-power / diagonal-gate circuits such as the hardware-efficient ansatz; see run_pass in sim_kernels.cuh).
#define QHBM_SWEEP_KERNELS_ONLY
#include "common.h"
#include "sim_launch.h"

namespace qhbm {

void launch_sweep_lean(int K, bool adj, bool dense, const KernelArgs& ka, unsigned grid, int threads, size_t smem,
                       cudaStream_t s) {
  launch_sweep_impl<false>(K, adj, dense, ka, grid, threads, smem, s);
}
void allow_large_smem_lean() { QHBM_CUDA(allow_large_smem_impl<false>()); }

}  // namespace qhbm
