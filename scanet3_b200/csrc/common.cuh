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
// ---- in-situ step timeline (compiled in only with -DSB_TRACE_BUILD: `make trace`; developer aid) --------------------------
// Every kernel of the engine passes griddepcontrol.wait exactly once; with tracing on, the first thread of its first CTA then
// appends %globaltimer to a device buffer (slot 0 = running count).  Inside a replayed CUDA graph this gives the start-to-start
// interval of consecutive kernels -- what each kernel costs the step WITH the overlap programmatic launches buy, which neither
// ncu (serialised, cold prologues) nor eager event timing (launch gaps) can show.  One __constant__ pointer per translation unit
// (no relocatable device code in this build).  NOT in the production library: even a null check per thread after the wait cost
// the issue-bound kernels of the LeNet step 2-5 % (1019 k -> 997 k -> 959 k samples/s with a wider predicate), so the product
// build compiles the stamp out and `make trace` builds scanet3_b200/libscanet_b200_trace.so for tools/trace_step.py.
static __constant__ unsigned long long* sb_trace_buf_c = nullptr;
static inline void sb_trace_set_local(unsigned long long* p) { cudaMemcpyToSymbol(sb_trace_buf_c, &p, sizeof p); }
constexpr int kTraceSlots = 1 << 16;
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef SB_TRACE_BUILD
  unsigned long long* tb = sb_trace_buf_c;
  if (tb != nullptr && (threadIdx.x | threadIdx.y | threadIdx.z | blockIdx.x | blockIdx.y | blockIdx.z) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned long long slot = 2 * atomicAdd(tb, 1ull) + 2;  // two words per event: time, (gridDim.x << 16 | blockDim.x)
    if (slot + 1 < (unsigned long long)kTraceSlots) {
      tb[slot] = t;
      tb[slot + 1] = ((unsigned long long)gridDim.x << 16) | blockDim.x;
    }
  }
#endif
}
__device__ __forceinline__ void pdl_prologue() { pdl_wait(); }
// Wait, then let the dependents launch: their CTAs fill whatever this grid leaves free and run their own pre-wait setup (a
// tensor-core kernel's barrier init, TMEM allocation, filter fill) under it; they still block in griddepcontrol.wait until this
// grid has completed.  Triggering AFTER the wait keeps the invariant the early filter fill relies on: when a kernel's pre-wait
// code runs, the grid two launches back is complete.  Used by the kernels of the CNN step, where it was measured kernel by kernel
// (cfg2: 854 k -> 906 k samples/s, profiles/r1_notes.md item 17); NOT by the generic fp32 GEMM / pointwise kernels: with it on
// every kernel the latency-bound MLP step (cfg1) lost 14 % (waiting dependents take slots from its multi-wave GEMMs).
__device__ __forceinline__ void pdl_prologue_trigger() {
  pdl_wait();
  pdl_trigger();
}
// the four kernels that the MLP step shares with the CNN step (prologue, optimizer, classifier-head loss, partial reduce) take the
// decision as an argument: the engine sets it per model (early_hint(): CNN path yes, plain fp32-GEMM MLP no)
__device__ __forceinline__ void pdl_prologue_trigger_if(int early) {
  pdl_wait();
  if (early) pdl_trigger();
}
bool pdl_enabled();  // SB_PDL=0 turns the launch attribute off (engine.cu)
int pdl_early_hint();               // host side, per thread: set by the engine for the model it is running (engine.cu)
void set_pdl_early_hint(int on);
// Pre-wait weight reads (a conv kernel's filter fill, the head kernels' W slab) are legal only if the kernel launched just before
// does not write weights: the dependency wait covers that kernel alone, and everything older is complete only because that
// kernel itself waited.  Host side, per thread: every launch sets it to 1, the launchers of weight-writing kernels (optimizer,
// initialisers, scaling, averaging) and every API entry set it to 0; kernels that read weights before their wait take it as an
// argument at launch (a forward pass started right behind an optimizer step -- the epoch's closing loss -- gets 0).
int prewait_ok();
void set_prewait_ok(int ok);
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
  set_prewait_ok(1);
  // developer aid (SB_SYNC_DEBUG=1, eager mode only): wait for every kernel and name the one that faults
  static const bool sync_debug = [] { const char* v = getenv("SB_SYNC_DEBUG"); return v && v[0] == '1'; }();
  if (sync_debug) {
    const cudaError_t err = cudaStreamSynchronize(s);
    if (err != cudaSuccess) {
      const char* name = "?";
      cudaFuncGetName(&name, (const void*)kernel);
      fprintf(stderr, "[sb] SB_SYNC_DEBUG: %s after kernel %s grid (%u,%u,%u) block (%u,%u,%u) smem %zu\n", cudaGetErrorString(err), name, grid.x,
              grid.y, grid.z, block.x, block.y, block.z, smem);
    }
  }
}
#endif

}  // namespace sb
