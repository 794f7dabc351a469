// Epoch-end synchronous model averaging across the k workers (GPUs) of one box.
// Replaces Spark's `broadcast` + `treeReduce(aggregateParams)` (optimizers/Optimizer.scala:72-86,196-230;
// models/Aggregation.scala:16-19): every tensor of the flat arena [model params | optimizer state] becomes a
// fixed convex combination of the k workers' tensors.
//
// SB_COLL_P2P: ONE kernel per GPU does reduce-scatter + all-gather over NVLink peer memory: rank r owns slice r
// of the arena, loads that slice from all k peers (P2P loads through NVSwitch), sums in the reduce's own order
// (so a chain of pairwise (l+r)/2 is reproduced bit for bit: power-of-two weights commute with fp32 rounding)
// and stores the result into slice r of all k arenas (P2P stores).  Each element is read and written by its
// owner rank only, so the reduction is in place with no intra-kernel cross-rank hazard.
// SB_COLL_NCCL: per-rank pre-scale + ncclAllReduce(sum); libnccl is loaded with dlopen so the library has no
// link-time NCCL dependency (the JVM / torch process brings its own).
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <memory>
#include <vector>

#include "common.cuh"
#include "engine.h"

using namespace sb;

namespace sb {
struct ApiError : std::runtime_error {
  int code;
  ApiError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void set_error(const std::string& m);
}  // namespace sb
#define SB_FAIL(code, msg) throw ::sb::ApiError((code), (msg))
#define SB_API_BEGIN try {
#define SB_API_END                                \
  }                                               \
  catch (const ::sb::ApiError& e) {               \
    ::sb::set_error(e.what());                    \
    return e.code;                                \
  }                                               \
  catch (const ::sb::CudaError& e) {              \
    ::sb::set_error(e.what());                    \
    return SB_ERR_CUDA;                           \
  }                                               \
  catch (const std::exception& e) {               \
    ::sb::set_error(e.what());                    \
    return SB_ERR_INVALID_ARG;                    \
  }                                               \
  return SB_OK;

constexpr int kMaxRanks = 8;

struct PeerSet {
  float* ptr[kMaxRanks];
  float w[kMaxRanks];
  int k;
  int n_groups;  // reduce shape: members of group j are the ranks q with q % n_groups == j, chained in ascending
                 // order; the group results are then chained in ascending j (Spark treeReduce)
};

// float4 slice [lo4, hi4) of the arena: load from every peer, weighted ordered sum, store to every peer
__global__ void __launch_bounds__(256) p2p_average_kernel(PeerSet ps, long long lo4, long long hi4) {
  pdl_prologue();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += stride) {
    float4 v[kMaxRanks];
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q)
      if (q < ps.k) v[q] = reinterpret_cast<const float4*>(ps.ptr[q])[i];  // all loads in flight before the math
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int gi = 0; gi < ps.n_groups; ++gi) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      bool first = true;
#pragma unroll
      for (int q = 0; q < kMaxRanks; ++q) {
        if (q >= ps.k || (q % ps.n_groups) != gi) continue;
        const float w = ps.w[q];
        float4 t = make_float4(__fmul_rn(v[q].x, w), __fmul_rn(v[q].y, w), __fmul_rn(v[q].z, w), __fmul_rn(v[q].w, w));
        if (first) { acc = t; first = false; }
        else { acc.x = __fadd_rn(acc.x, t.x); acc.y = __fadd_rn(acc.y, t.y); acc.z = __fadd_rn(acc.z, t.z); acc.w = __fadd_rn(acc.w, t.w); }
      }
      if (gi == 0) tot = acc;
      else { tot.x = __fadd_rn(tot.x, acc.x); tot.y = __fadd_rn(tot.y, acc.y); tot.z = __fadd_rn(tot.z, acc.z); tot.w = __fadd_rn(tot.w, acc.w); }
    }
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q)
      if (q < ps.k) reinterpret_cast<float4*>(ps.ptr[q])[i] = tot;
  }
}

// ---- device-synchronised variant: flag barrier + reduce + loss combine + flag barrier in ONE kernel per rank ----
// Tail of every arena allocation (kArenaTailFloats floats behind the tensors), written by peers over NVLink:
//   u64 arrive[8] | u64 done[8] | float loss[8] | (pad to 64 floats) | i64 iters[8]        (slot q is written by rank q)
// The tail is zeroed by sb_group_create_ipc (flags of an earlier group on the same engine must not satisfy this group's
// barriers); the host plumbing puts a cross-process barrier between creation and first use.
constexpr int kTailLossOff = 32;   // floats
constexpr int kTailItersOff = 64;  // floats (8-byte aligned: the arena is padded to 32 floats)
struct SyncArgs {
  long long tail_off;  // floats from the arena base
  unsigned long long epoch;
  int rank;
  StepState* st;        // this engine's step state: loss read before the barrier and replaced by the combined loss; iter gets
                        // the peers' iteration counts added (Optimizer.scala:88,223); err receives the timeout flag
  long long my_iters;   // iterations this rank ran in the epoch that just ended
  unsigned int* ticket; // local CTA-completion counter (zero between launches)
  unsigned long long timeout_ns;
};
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded in wall-clock time (SB_P2P_TIMEOUT_MS, default 10 min): a rank that never arrives makes this return false; the
// kernel then records the failure in StepState.err and exits WITHOUT reducing -- the context stays usable and the host gets
// SB_ERR_TIMEOUT at its next sync point (sb_sync / sb_read_loss) instead of a sticky launch failure.
__device__ __forceinline__ bool wait_flag(const unsigned long long* p, unsigned long long epoch, unsigned long long t0,
                                          unsigned long long timeout_ns) {
  unsigned spins = 0;
  while (ld_acquire_sys(p) < epoch) {
    __nanosleep(spins < 64 ? 100 : 1000);
    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) return false;
  }
  return true;
}
__global__ void __launch_bounds__(256) p2p_average_sync_kernel(PeerSet ps, long long lo4, long long hi4, SyncArgs a) {
  pdl_prologue();
  __shared__ int bail;
  unsigned long long* my_tail = reinterpret_cast<unsigned long long*>(ps.ptr[a.rank] + a.tail_off);
  if (threadIdx.x == 0) {
    const unsigned long long t0 = globaltimer_ns();
    bool ok = true;
    if (blockIdx.x == 0) {
      const float l = a.st->loss;
      for (int q = 0; q < ps.k; ++q) {
        (ps.ptr[q] + a.tail_off + kTailLossOff)[a.rank] = l;
        reinterpret_cast<long long*>(ps.ptr[q] + a.tail_off + kTailItersOff)[a.rank] = a.my_iters;
      }
      __threadfence_system();
      for (int q = 0; q < ps.k; ++q)
        st_release_sys(reinterpret_cast<unsigned long long*>(ps.ptr[q] + a.tail_off) + a.rank, a.epoch);
    }
    for (int q = 0; q < ps.k && ok; ++q) ok = wait_flag(my_tail + q, a.epoch, t0, a.timeout_ns);  // every rank finished its epoch
    if (ok && blockIdx.x == 0) {
      // loss: the same weights in the same order (Optimizer.scala:221); iterations: a plain sum (Optimizer.scala:223)
      const float* ls = ps.ptr[a.rank] + a.tail_off + kTailLossOff;
      const long long* its = reinterpret_cast<const long long*>(ps.ptr[a.rank] + a.tail_off + kTailItersOff);
      float tot = 0.f;
      for (int gi = 0; gi < ps.n_groups; ++gi) {
        float acc = 0.f;
        bool first = true;
        for (int q = gi; q < ps.k; q += ps.n_groups) {
          const float t = __fmul_rn(__ldcv(ls + q), ps.w[q]);
          acc = first ? t : __fadd_rn(acc, t);
          first = false;
        }
        tot = gi == 0 ? acc : __fadd_rn(tot, acc);
      }
      long long others = 0;
      for (int q = 0; q < ps.k; ++q)
        if (q != a.rank) others += __ldcv(its + q);
      a.st->loss = tot;
      a.st->avg_total_iters = others + a.my_iters;
      a.st->iter += others;  // the step counter is global: next epoch starts at globalIter + sum of all ranks' counts
    }
    if (!ok) a.st->err = 1;
    bail = ok ? 0 : 1;
  }
  __syncthreads();
  if (bail) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += stride) {
    float4 v[kMaxRanks];
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q)
      if (q < ps.k) v[q] = __ldcv(reinterpret_cast<const float4*>(ps.ptr[q]) + i);
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int gi = 0; gi < ps.n_groups; ++gi) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      bool first = true;
#pragma unroll
      for (int q = 0; q < kMaxRanks; ++q) {
        if (q >= ps.k || (q % ps.n_groups) != gi) continue;
        const float w = ps.w[q];
        float4 t = make_float4(__fmul_rn(v[q].x, w), __fmul_rn(v[q].y, w), __fmul_rn(v[q].z, w), __fmul_rn(v[q].w, w));
        if (first) { acc = t; first = false; }
        else { acc.x = __fadd_rn(acc.x, t.x); acc.y = __fadd_rn(acc.y, t.y); acc.z = __fadd_rn(acc.z, t.z); acc.w = __fadd_rn(acc.w, t.w); }
      }
      if (gi == 0) tot = acc;
      else { tot.x = __fadd_rn(tot.x, acc.x); tot.y = __fadd_rn(tot.y, acc.y); tot.z = __fadd_rn(tot.z, acc.z); tot.w = __fadd_rn(tot.w, acc.w); }
    }
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q)
      if (q < ps.k) reinterpret_cast<float4*>(ps.ptr[q])[i] = tot;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(a.ticket, 1u) == gridDim.x - 1) {  // last CTA of this rank: its slice is stored everywhere
      *a.ticket = 0;
      const unsigned long long t0 = globaltimer_ns();
      for (int q = 0; q < ps.k; ++q)
        st_release_sys(reinterpret_cast<unsigned long long*>(ps.ptr[q] + a.tail_off) + 8 + a.rank, a.epoch);
      bool ok = true;
      for (int q = 0; q < ps.k && ok; ++q) ok = wait_flag(my_tail + 8 + q, a.epoch, t0, a.timeout_ns);  // every slice of MY arena has landed
      if (!ok) a.st->err = 1;
    }
  }
}

// ---- Spark 3.3.1 RDD.treeReduce(f, depth = 2) shape under ascending completion order ----
static void chain_weights(int n, std::vector<double>& w) {
  w.assign(n, 1.0);
  if (n == 1) return;
  w[0] = pow(0.5, n - 1);
  for (int i = 1; i < n; ++i) w[i] = pow(0.5, n - i);
}
static int spark_shape(int k, std::vector<double>& weights) {
  // returns the number of first-level groups (1 = a single chain)
  int parts = k;
  int scale = std::max((int)ceil(sqrt((double)parts)), 2);
  int n_groups = 1;
  std::vector<double> w(k, 1.0);
  std::vector<std::vector<int>> groups(k);
  for (int i = 0; i < k; ++i) groups[i] = {i};
  while (parts > scale + (int)ceil((double)parts / scale)) {
    parts /= scale;
    std::vector<std::vector<int>> merged(parts);
    std::vector<std::vector<std::vector<int>>> buckets(parts);
    for (size_t i = 0; i < groups.size(); ++i) buckets[i % parts].push_back(groups[i]);
    for (int b = 0; b < parts; ++b) {
      std::vector<double> cw;
      chain_weights((int)buckets[b].size(), cw);
      for (size_t j = 0; j < buckets[b].size(); ++j)
        for (int idx : buckets[b][j]) {
          w[idx] *= cw[j];
          merged[b].push_back(idx);
        }
    }
    groups = merged;
    n_groups = parts;
  }
  std::vector<double> cw;
  chain_weights((int)groups.size(), cw);
  for (size_t j = 0; j < groups.size(); ++j)
    for (int idx : groups[j]) w[idx] *= cw[j];
  weights = w;
  return n_groups;
}

extern "C" int sb_spark_chain_weights(int k, float* weights_out) {
  SB_API_BEGIN
  if (k < 1 || k > 64 || !weights_out) SB_FAIL(SB_ERR_INVALID_ARG, "k must be in [1,64]");
  std::vector<double> w;
  spark_shape(k, w);
  for (int i = 0; i < k; ++i) weights_out[i] = (float)w[i];
  SB_API_END
}

// ---- NCCL through dlopen ----
typedef struct ncclComm* ncclComm_t;
struct NcclApi {
  void* lib = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static bool load_nccl(NcclApi& api, std::string& err) {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) {
    err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    return false;
  }
  api.CommInitAll = (int (*)(ncclComm_t*, int, const int*))dlsym(api.lib, "ncclCommInitAll");
  api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.lib, "ncclCommDestroy");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
  api.GroupStart = (int (*)())dlsym(api.lib, "ncclGroupStart");
  api.GroupEnd = (int (*)())dlsym(api.lib, "ncclGroupEnd");
  api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
  if (!api.CommInitAll || !api.CommDestroy || !api.AllReduce || !api.GroupStart || !api.GroupEnd) {
    err = "libnccl is missing required symbols";
    return false;
  }
  return true;
}

struct sb_group {
  std::vector<sb_engine*> es;  // single-process mode: all k engines; IPC mode: only this rank's engine at [rank]
  int k = 0;
  int rank = -1;  // >= 0 in IPC (one process per GPU) mode
  int collective = SB_COLL_P2P;
  std::vector<float> w;
  int n_groups = 1;
  PeerSet ps[kMaxRanks];  // per local rank: peer pointers as seen from that device
  std::vector<void*> ipc_opened;
  unsigned int* ticket = nullptr;  // device counter of the device-synchronised kernel
  uint64_t last_epoch = 0;         // sb_group_average_device: epochs must increase
  NcclApi nccl;
  std::vector<ncclComm_t> comms;
};

static void fill_weights(sb_group* g, int avg_mode, const float* weights) {
  const int k = g->k;
  g->w.assign(k, 1.0f / k);
  g->n_groups = 1;
  if (avg_mode == SB_AVG_SPARK_CHAIN) {
    std::vector<double> w;
    g->n_groups = spark_shape(k, w);
    for (int i = 0; i < k; ++i) g->w[i] = (float)w[i];
  } else if (avg_mode == SB_AVG_WEIGHTS) {
    if (!weights) SB_FAIL(SB_ERR_INVALID_ARG, "SB_AVG_WEIGHTS needs a weight vector");
    for (int i = 0; i < k; ++i) g->w[i] = weights[i];
  } else if (avg_mode != SB_AVG_MEAN) {
    SB_FAIL(SB_ERR_INVALID_ARG, "unknown avg_mode");
  }
}

static void check_same_arena(const std::vector<sb_engine*>& es) {
  for (sb_engine* e : es) {
    if (!e || e->plan_only) SB_FAIL(SB_ERR_CUDA, "group members must be device engines");
    if (e->arena_floats != es[0]->arena_floats || e->params.size() != es[0]->params.size())
      SB_FAIL(SB_ERR_INVALID_ARG, "group members must share one model/optimizer layout");
  }
}

extern "C" int sb_group_create(sb_engine** engines, int k, int avg_mode, const float* weights_or_null, int collective,
                               sb_group** out) {
  SB_API_BEGIN
  if (!engines || !out || k < 1 || k > kMaxRanks) SB_FAIL(SB_ERR_INVALID_ARG, "k must be in [1,8]");
  std::unique_ptr<sb_group> g(new sb_group());
  g->k = k;
  g->es.assign(engines, engines + k);
  g->collective = collective;
  check_same_arena(g->es);
  fill_weights(g.get(), avg_mode, weights_or_null);
  if (collective == SB_COLL_NCCL && g->n_groups > 1)
    SB_FAIL(SB_ERR_UNSUPPORTED, "SB_COLL_NCCL sums in NCCL's own order and cannot reproduce the grouped Spark chain (k >= 6): use "
                                "SB_COLL_P2P, or SB_AVG_MEAN / SB_AVG_WEIGHTS with NCCL");
  if (collective == SB_COLL_P2P) {
    for (int r = 0; r < k; ++r) {
      SB_CUDA(cudaSetDevice(g->es[r]->device));
      for (int q = 0; q < k; ++q) {
        if (g->es[q]->device == g->es[r]->device) continue;
        int can = 0;
        SB_CUDA(cudaDeviceCanAccessPeer(&can, g->es[r]->device, g->es[q]->device));
        if (!can) SB_FAIL(SB_ERR_UNSUPPORTED, "no peer access between the group's devices; use SB_COLL_NCCL");
        cudaError_t ce = cudaDeviceEnablePeerAccess(g->es[q]->device, 0);
        if (ce == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else SB_CUDA(ce);
      }
      PeerSet& ps = g->ps[r];
      ps.k = k;
      ps.n_groups = g->n_groups;
      for (int q = 0; q < k; ++q) { ps.ptr[q] = g->es[q]->arena; ps.w[q] = g->w[q]; }
    }
  } else if (collective == SB_COLL_NCCL) {
    std::string err;
    if (!load_nccl(g->nccl, err)) SB_FAIL(SB_ERR_NCCL, err);
    std::vector<int> devs(k);
    for (int r = 0; r < k; ++r) devs[r] = g->es[r]->device;
    g->comms.resize(k);
    int rc = g->nccl.CommInitAll(g->comms.data(), k, devs.data());
    if (rc != 0) SB_FAIL(SB_ERR_NCCL, std::string("ncclCommInitAll: ") + (g->nccl.GetErrorString ? g->nccl.GetErrorString(rc) : "error"));
  } else {
    SB_FAIL(SB_ERR_INVALID_ARG, "unknown collective");
  }
  *out = g.release();
  SB_API_END
}

// one process per GPU: peers' arenas arrive as CUDA IPC handles (exchanged by the host plumbing, e.g.
// torch.distributed.all_gather_object); barriers around sb_group_average_local are the caller's.
extern "C" int sb_group_create_ipc(sb_engine* self, int rank, int k, const void* handles, int avg_mode,
                                   const float* weights_or_null, sb_group** out) {
  SB_API_BEGIN
  if (!self || self->plan_only) SB_FAIL(SB_ERR_CUDA, "device engine required");
  if (!handles || !out || k < 1 || k > kMaxRanks || rank < 0 || rank >= k) SB_FAIL(SB_ERR_INVALID_ARG, "bad rank/k");
  std::unique_ptr<sb_group> g(new sb_group());
  g->k = k;
  g->rank = rank;
  g->es.assign(k, nullptr);
  g->es[rank] = self;
  g->collective = SB_COLL_P2P;
  fill_weights(g.get(), avg_mode, weights_or_null);
  SB_CUDA(cudaSetDevice(self->device));
  PeerSet& ps = g->ps[rank];
  ps.k = k;
  ps.n_groups = g->n_groups;
  for (int q = 0; q < k; ++q) {
    ps.w[q] = g->w[q];
    if (q == rank) { ps.ptr[q] = self->arena; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)q * 64, 64);
    void* p = nullptr;
    SB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    g->ipc_opened.push_back(p);
    ps.ptr[q] = (float*)p;
  }
  // flags / scalars an earlier group left in this rank's tail must not satisfy this group's barriers (epochs restart at 1):
  // zero it now; the caller barriers across the ranks between this call and the first sb_group_average_device
  SB_CUDA(cudaMemsetAsync(self->arena + self->arena_floats, 0, (size_t)kArenaTailFloats * 4, self->stream));
  SB_CUDA(cudaMalloc((void**)&g->ticket, 64));
  SB_CUDA(cudaMemsetAsync(g->ticket, 0, 64, self->stream));
  SB_CUDA(cudaStreamSynchronize(self->stream));
  *out = g.release();
  SB_API_END
}

static void launch_slice(sb_group* g, int r) {
  sb_engine* e = g->es[r];
  SB_CUDA(cudaSetDevice(e->device));
  const long long n4 = e->arena_floats / 4;  // tensors are padded to 32 floats
  const long long per = (n4 + g->k - 1) / g->k;
  const long long lo = per * r, hi = std::min(n4, per * (r + 1));
  if (hi <= lo) return;
  int blocks = (int)std::min<long long>((hi - lo + 255) / 256, (long long)kNumSMs * 4);
  launch_k(p2p_average_kernel, blocks, 256, 0, e->stream, g->ps[r], lo, hi);
  SB_LAUNCHED();
  set_prewait_ok(0);  // writes weights
  e->launches += 1;
}

extern "C" int sb_group_average_local(sb_group* g) {
  SB_API_BEGIN
  if (!g || g->rank < 0) SB_FAIL(SB_ERR_INVALID_ARG, "not an IPC group");
  if (g->k > 1) launch_slice(g, g->rank);
  SB_API_END
}

static unsigned long long p2p_timeout_ns() {
  static const unsigned long long ns = [] {
    const char* v = getenv("SB_P2P_TIMEOUT_MS");
    const double ms = v ? atof(v) : 600000.0;
    return (unsigned long long)((ms > 1.0 ? ms : 1.0) * 1e6);
  }();
  return ns;
}

extern "C" int sb_group_average_device(sb_group* g, uint64_t epoch, int64_t my_iterations) {
  SB_API_BEGIN
  if (!g || g->rank < 0) SB_FAIL(SB_ERR_INVALID_ARG, "not an IPC group");
  if (epoch <= g->last_epoch)
    SB_FAIL(SB_ERR_INVALID_ARG, "sb_group_average_device: epoch " + std::to_string(epoch) + " does not exceed the last one used (" +
                                    std::to_string(g->last_epoch) + "): stale flags would open the barrier");
  g->last_epoch = epoch;
  if (g->k == 1) return SB_OK;
  sb_engine* e = g->es[g->rank];
  SB_CUDA(cudaSetDevice(e->device));
  e->device_sync_used = true;
  e->host_iter_mirror = -1;  // the kernel adds the peers' iteration counts to the device counter
  SB_CUDA(cudaMemsetAsync(g->ticket, 0, 4, e->stream));  // a timed-out launch leaves it non-zero
  const long long n4 = e->arena_floats / 4;
  const long long per = (n4 + g->k - 1) / g->k;
  const long long lo = per * g->rank, hi = std::min(n4, per * (g->rank + 1));
  // every CTA polls flags, so the whole grid must be co-resident: <= 8 CTAs of 256 threads per SM (SB_P2P_CTAS_PER_SM tunes)
  static const int per_sm = [] { const char* v = getenv("SB_P2P_CTAS_PER_SM"); int k = v ? atoi(v) : 4; return k < 1 ? 1 : (k > 8 ? 8 : k); }();
  int blocks = (int)std::min<long long>(std::max<long long>((hi - lo + 255) / 256, 1), (long long)kNumSMs * per_sm);
  SyncArgs a;
  a.tail_off = e->arena_floats;
  a.epoch = epoch;
  a.rank = g->rank;
  a.st = e->st;
  a.my_iters = my_iterations;
  a.ticket = g->ticket;
  a.timeout_ns = p2p_timeout_ns();
  launch_k(p2p_average_sync_kernel, blocks, 256, 0, e->stream, g->ps[g->rank], lo, hi, a);
  SB_LAUNCHED();
  set_prewait_ok(0);  // writes weights
  e->launches += 1;
  int rc = sb_reset_unaggregated(e);  // aggregation == None tensors are re-initialised (Optimizer.scala:209)
  if (rc != SB_OK) return rc;
  SB_API_END
}

extern "C" int sb_group_read_result(sb_group* g, float* loss_out, int64_t* total_iterations_out) {
  SB_API_BEGIN
  if (!g || g->rank < 0) SB_FAIL(SB_ERR_INVALID_ARG, "not an IPC group");
  sb_engine* e = g->es[g->rank];
  SB_CUDA(cudaSetDevice(e->device));
  float loss = 0.f;
  int rc = sb_read_loss(e, &loss);  // syncs the stream; SB_ERR_TIMEOUT if a peer never reached the barrier
  if (rc != SB_OK) return rc;
  if (loss_out) *loss_out = loss;
  if (total_iterations_out) {
    long long t = 0;
    SB_CUDA(cudaMemcpy(&t, &e->st->avg_total_iters, sizeof t, cudaMemcpyDeviceToHost));
    *total_iterations_out = t;
  }
  SB_API_END
}

extern "C" int sb_group_average(sb_group* g, float* loss_inout, int64_t* iterations_inout) {
  SB_API_BEGIN
  if (!g || g->rank >= 0) SB_FAIL(SB_ERR_INVALID_ARG, "single-process group required");
  const int k = g->k;
  // loss combines with the same weights in the same order; iterations are a plain sum (Optimizer.scala:221-223)
  if (loss_inout) {
    float tot = 0.f;
    for (int gi = 0; gi < g->n_groups; ++gi) {
      float acc = 0.f;
      bool first = true;
      for (int q = gi; q < k; q += g->n_groups) {
        float t = loss_inout[q] * g->w[q];
        acc = first ? t : acc + t;
        first = false;
      }
      tot = gi == 0 ? acc : tot + acc;
    }
    for (int q = 0; q < k; ++q) loss_inout[q] = tot;
  }
  if (iterations_inout) {
    int64_t s = 0;
    for (int q = 0; q < k; ++q) s += iterations_inout[q];
    for (int q = 0; q < k; ++q) iterations_inout[q] = s;
  }
  if (k == 1) return SB_OK;  // no combine at all (aggregateParams never runs)
  for (int r = 0; r < k; ++r) {  // every worker finished its epoch
    SB_CUDA(cudaSetDevice(g->es[r]->device));
    SB_CUDA(cudaStreamSynchronize(g->es[r]->stream));
  }
  if (g->collective == SB_COLL_P2P) {
    for (int r = 0; r < k; ++r) launch_slice(g, r);
  } else {
    for (int r = 0; r < k; ++r) {
      sb_engine* e = g->es[r];
      SB_CUDA(cudaSetDevice(e->device));
      scale_inplace(e->arena, g->w[r], e->arena_floats, e->stream);
      e->launches += 1;
    }
    g->nccl.GroupStart();
    for (int r = 0; r < k; ++r) {
      sb_engine* e = g->es[r];
      int rc = g->nccl.AllReduce(e->arena, e->arena, (size_t)e->arena_floats, /*ncclFloat32*/ 7, /*ncclSum*/ 0, g->comms[r], e->stream);
      if (rc != 0) {
        g->nccl.GroupEnd();
        SB_FAIL(SB_ERR_NCCL, "ncclAllReduce failed");
      }
    }
    int rc = g->nccl.GroupEnd();
    if (rc != 0) SB_FAIL(SB_ERR_NCCL, "ncclGroupEnd failed");
  }
  for (int r = 0; r < k; ++r) {
    SB_CUDA(cudaSetDevice(g->es[r]->device));
    SB_CUDA(cudaStreamSynchronize(g->es[r]->stream));
  }
  for (int r = 0; r < k; ++r) {
    int rc = sb_reset_unaggregated(g->es[r]);  // aggregation == None tensors are re-initialised (Optimizer.scala:209)
    if (rc != SB_OK) return rc;
  }
  SB_API_END
}

extern "C" int sb_group_weights(const sb_group* g, float* weights_out) {
  SB_API_BEGIN
  if (!g || !weights_out) SB_FAIL(SB_ERR_INVALID_ARG, "null argument");
  for (int i = 0; i < g->k; ++i) weights_out[i] = g->w[i];
  SB_API_END
}

extern "C" void sb_group_destroy(sb_group* g) {
  if (!g) return;
  for (void* p : g->ipc_opened) cudaIpcCloseMemHandle(p);
  if (g->ticket) cudaFree(g->ticket);
  if (g->nccl.CommDestroy)
    for (ncclComm_t c : g->comms)
      if (c) g->nccl.CommDestroy(c);
  delete g;
}

namespace sb {
void trace_set_group(unsigned long long* p) { sb_trace_set_local(p); }  // SB_TRACE timeline buffer of this translation unit
}  // namespace sb
