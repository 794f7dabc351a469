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

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// Every kernel of the step calls pdl_wait() (griddepcontrol.wait) before it touches global memory: it returns once the
// PREVIOUS kernel of the stream has completed and flushed, so memory ordering is exactly that of plain stream order, while the
// kernel's CTAs may already be scheduled when the previous kernel's CTAs have exited.  Simple kernels wait first thing
// (pdl_prologue); the tensor-core kernels run their on-chip setup (mbarrier init, TMEM allocation, tensor-map prefetch) first.
// The instruction is a no-op when a kernel is launched without the PDL attribute (SB_PDL=0).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
#ifdef SB_PDL_EARLY_TRIGGER
  pdl_trigger();  // measured on cfg2: resident-but-waiting CTAs of the next kernel slow multi-wave kernels down (-2.5 %)
#endif
  pdl_wait();
}
bool pdl_enabled();  // SB_PDL=0 turns the launch attribute off (engine.cu)
template <typename K, typename... A>
inline void launch_k(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SB_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<A&&>(args)...));
}
#endif

}  // namespace sb
