// common.h -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

namespace qhbm {

void set_last_error(const std::string& msg);

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess)
    throw CudaError(std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define QHBM_CUDA(x) ::qhbm::cuda_check((x), #x)

// Runs `body`, converts exceptions into a status code + qhbm_last_error().
template <class F>
int guarded(F&& body) {
  try {
    body();
    return 0;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return 1;
  } catch (...) {
    set_last_error("unknown error");
    return 1;
  }
}

}  // namespace qhbm
