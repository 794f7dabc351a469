#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>

namespace sb {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    throw CudaError(buf);
  }
}
#define SB_CUDA(x) ::sb::cuda_check((x), #x, __FILE__, __LINE__)

// counted launch check: every launcher ends with SB_LAUNCHED()
extern thread_local long long tl_launch_count;
#define SB_LAUNCHED()                      \
  do {                                     \
    ++::sb::tl_launch_count;               \
    SB_CUDA(cudaPeekAtLastError());        \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace sb
