// Persistent LSTM time loop (forward) on tcgen05: ONE launch walks all T steps of RNN(LSTMCell) (layer/RNN.scala:41-73, 319-339)
// instead of one GEMM + one gate kernel per step (cfg3: 131 launches, ~10.8 us per step of which ~1 us is math).
//
//   CTA (mi, ni) owns 128 batch rows x 64 gate columns (= 16 units x 4 gates, unit-interleaved as in lstm.cu: column 4*j + G).
//   Its slice of Wr_cat [U x 64] stays RESIDENT in shared memory for the whole sequence; per step the 128 x U block of h_{t-1}
//   streams in by TMA (K-major, 32-wide chunks through a 4-stage ring), U/8 MMAs (M = 128, N = 64, K = 8, kind::tf32) accumulate
//   Z_t in TMEM, and eight epilogue warps turn it into the step's result: z + x_t.Wk + b -> activations -> acts[t] (saved for
//   BPTT), c_t (kept in REGISTERS across the steps, also stored: BPTT needs every c_t), h_t -> h[t+1].
//   Step t+1 needs the whole row block of h_t: the NB column-CTAs of a row block meet at a counter in global memory
//   (release add after the h stores; the TMA producer polls it with acquire loads and a generic->async proxy fence before it
//   issues the loads).  All B/128 x 4U/64 CTAs (128 for cfg3) are co-resident (one per SM; the launcher refuses grids > 148),
//   every wait is bounded (trap instead of a hang).
//   Numerics: the same expressions, in the same order, as lstm_gates_fwd_kernel<.., true> (TF32 mode); the recurrent product in TF32.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "engine.h"
#include "tc.h"
#include "tc_sm100.cuh"

namespace sb {

using namespace tc;
CUtensorMap make_tma_map(const float* base, int rank, const uint64_t* dims, const uint32_t* box, bool mn_major);  // tc.cu

namespace {

constexpr int kLpThreads = 320;   // warp 0: TMA + row-block barrier, warp 1: MMA issuer + TMEM owner, warps 2..9: epilogue
constexpr int kLpStages = 4;
constexpr int kLpATile = 128 * 128;  // 128 rows x 32 k floats
constexpr int kLpMaxF = 8;

struct LstmPersistParams {
  int B, T, U, F, NB;      // NB = 4U / 64 column blocks per row block
  int act, ract, has_bias;
  int zero_state;          // stateless: h_{-1} = c_{-1} = 0, step 0 has no recurrent product
  const float* x_tm;       // [T, B, F]
  const float* wk_cat;     // [F, 4U]
  const float* b_cat;      // [4U]
  const float* c_init;     // [B, U] or null
  float* acts;             // [T, B, 4U]
  float* c;                // [T, B, U]
  float* h;                // [T+1, B, U]
  unsigned int* bar;       // [B / 128] arrival counters, zero at launch
  unsigned long long* dbg; // developer timeline of the cluster form (SB_LC_TIMELINE=1): [T][8] %globaltimer stamps of CTA 0, else null
};

// the TF32-mode forms of lstm.cu actf_t<true> (ex2.approx / rcp.approx): these kernels run in TF32 mode only
__device__ __forceinline__ float lp_act(float x, int act) {
  switch (act) {
    case 1: return __fdividef(1.f, 1.f + __expf(-x));
    case 2: return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x));
    case 4: return fmaxf(0.f, x);
    default: return x;
  }
}
// h is stored rounded to the nearest tf32 value, as lstm_gates_fwd_kernel<.., true> does (lstm.cu tf32_rn): these kernels run in TF32 mode only
__device__ __forceinline__ float lp_tf32_rn(float x) {
  unsigned u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kLpThreads, 1)
lstm_persist_fwd_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_wr, const LstmPersistParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int kchunks = p.U >> 5;
  const int b_bytes = kchunks * 2 * 4096;                     // resident Wr slice: per 32-k chunk two [32 n x 32 k] boxes
  uint8_t* sm_b = smem;
  uint8_t* sm_a = smem + b_bytes;
  float* s_wk = reinterpret_cast<float*>(sm_a + kLpStages * kLpATile);   // [F][64] input-kernel slice
  float* s_b = s_wk + kLpMaxF * 64;                                       // [64]   bias slice
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_b + 64);
  uint64_t* empty_bar = full_bar + kLpStages;
  uint64_t* tmem_full = empty_bar + kLpStages;
  uint64_t* tmem_empty = tmem_full + 1;
  uint64_t* b_bar = tmem_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int mi = blockIdx.x / p.NB, ni = blockIdx.x - mi * p.NB;
  const int col0 = ni * 64, row0 = mi * 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_h);
    tma_prefetch_desc(&map_wr);
    for (int s = 0; s < kLpStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 8);
    mbar_init(b_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the pack kernels before this one wrote Wr_cat / Wk_cat / b_cat
  for (int i = threadIdx.x; i < p.F * 64; i += blockDim.x) s_wk[i] = p.wk_cat[(size_t)(i >> 6) * 4 * p.U + col0 + (i & 63)];
  if (threadIdx.x < 64) s_b[threadIdx.x] = p.has_bias ? p.b_cat[col0 + threadIdx.x] : 0.f;
  __syncthreads();
  const int t_first = p.zero_state ? 1 : 0;  // first step with a recurrent product

  if (warp == 0) {
    // ================= TMA producer (+ the row block's step barrier) =================
    if (lane == 0) {
      mbar_arrive_expect_tx(b_bar, (uint32_t)b_bytes);
      for (int kc = 0; kc < kchunks; ++kc)
        for (int j = 0; j < 2; ++j) tma_load_2d(sm_b + (size_t)(kc * 2 + j) * 4096, &map_wr, b_bar, col0 + j * 32, kc * 32);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_first; t < p.T; ++t) {
        if (t > 0) {  // every column block of this row block has stored its slice of h_{t-1}
          const unsigned int want = (unsigned int)(p.NB * t);
          unsigned long long spins = 0;
          while (ld_acquire_gpu(p.bar + mi) < want) {
            __nanosleep(32);
            if (++spins > (1ull << 26)) __trap();
          }
          asm volatile("fence.proxy.async;" ::: "memory");  // peers' generic-proxy stores -> this thread's async-proxy (TMA) loads
        }
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)kLpATile);
          tma_load_2d(sm_a + (size_t)stage * kLpATile, &map_h, &full_bar[stage], kc * 32, t * p.B + row0);  // h slot t = h_{t-1}
          if (++stage == kLpStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_tf32(128, 64, 0, 1);
    const uint64_t a_d = smem_desc_sw128(0, 16, 1024), b_d = smem_desc(0, 4096, 512, 1);
    const uint32_t a_hi = (uint32_t)(a_d >> 32), a_lo0 = (uint32_t)a_d, b_hi = (uint32_t)(b_d >> 32), b_lo0 = (uint32_t)b_d;
    const uint32_t sb0 = b_lo0 + (smem_u32(sm_b) >> 4);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    mbar_wait(b_bar, 0);
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0, eph = 0;
    for (int t = t_first; t < p.T; ++t) {
      mbar_wait(tmem_empty, eph ^ 1);  // the epilogue has drained the previous step's accumulator
      eph ^= 1;
      tc_fence_after();
      uint32_t accumulate = 0;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t at = a_lo0 + (smem_u32(sm_a + (size_t)stage * kLpATile) >> 4);
        const uint32_t bt = sb0 + (uint32_t)(kc * (8192 >> 4));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_tf32_elect(tmem_u, ((uint64_t)a_hi << 32) | (at + ks * 2), ((uint64_t)b_hi << 32) | (bt + ks * (1024 >> 4)), idesc, accumulate);
          accumulate = 1;
        }
        umma_commit_elect(&empty_bar[stage]);
        if (++stage == kLpStages) { stage = 0; phase ^= 1; }
      }
      umma_commit_elect(tmem_full);
    }
  } else {
    // ================= epilogue: gates, c_t (registers), h_t; eight warps: sub-partition x column half =================
    const int sub = warp & 3, half = (warp - 2) >> 2;
    const int r = sub * 32 + lane, row = row0 + r;
    const int cbase = half * 32;              // this thread's 32 gate columns inside the CTA's 64 = 8 units
    const int u0 = (col0 + cbase) >> 2;       // first of its 8 units
    float cst[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cst[j] = p.c_init ? p.c_init[(size_t)row * p.U + u0 + j] : 0.f;
    uint32_t fph = 0;
    for (int t = 0; t < p.T; ++t) {
      float z[32];
      const bool has_z = t >= t_first;
      if (has_z) {
        mbar_wait(tmem_full, fph);
        fph ^= 1;
        tc_fence_after();
        tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)cbase, z);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty);  // drained: the next step's MMAs may overwrite it
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) z[q] = 0.f;
      }
      float xv[kLpMaxF];
#pragma unroll
      for (int f = 0; f < kLpMaxF; ++f) xv[f] = f < p.F ? __ldg(p.x_tm + ((size_t)t * p.B + row) * p.F + f) : 0.f;
      float* arow = p.acts + ((size_t)t * p.B + row) * 4 * p.U + col0 + cbase;
      float hq[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cc = cbase + 4 * j + g;
          float xp = 0.f;
#pragma unroll
          for (int f = 0; f < kLpMaxF; ++f)
            if (f < p.F) xp = fmaf(xv[f], s_wk[f * 64 + cc], xp);
          float v = has_z ? __fadd_rn(xp, z[4 * j + g]) : xp;       // (x.Wk + h.Wr)
          if (p.has_bias) v = __fadd_rn(v, s_b[cc]);                  // + b
          a[g] = lp_act(v, g == 2 ? p.act : p.ract);
        }
        *reinterpret_cast<float4*>(arow + 4 * j) = make_float4(a[0], a[1], a[2], a[3]);
        const float cn = __fadd_rn(__fmul_rn(cst[j], a[0]), __fmul_rn(a[1], a[2]));   // c_prev*f + i*g
        cst[j] = cn;
        hq[j] = lp_tf32_rn(__fmul_rn(a[3], lp_act(cn, p.act)));                                    // o * act(c)
      }
      float* crow = p.c + ((size_t)t * p.B + row) * p.U + u0;
      float* hrow = p.h + ((size_t)(t + 1) * p.B + row) * p.U + u0;
      *reinterpret_cast<float4*>(crow) = make_float4(cst[0], cst[1], cst[2], cst[3]);
      *reinterpret_cast<float4*>(crow + 4) = make_float4(cst[4], cst[5], cst[6], cst[7]);
      *reinterpret_cast<float4*>(hrow) = make_float4(hq[0], hq[1], hq[2], hq[3]);
      *reinterpret_cast<float4*>(hrow + 4) = make_float4(hq[4], hq[5], hq[6], hq[7]);
      // this CTA's slice of h_t is complete once all eight epilogue warps are here: one release-add for the row block
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 64) red_release_gpu_add(p.bar + mi, 1u);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// =====================================================================================================================
// Cluster form (SB_LSTM_PERSIST=2): 8-CTA clusters, no global-memory synchronisation on the step-to-step path.
//   Cluster = one block of 64 batch rows; CTA `rank` owns 128 gate columns (= 32 units = ONE 32-wide K chunk of the next step's A
//   operand) with its slice of Wr_cat [U x 128] resident (128 KB) and the whole 64 x U block of h_{t-1} in a single A buffer
//   (U/32 chunks of 8 KB).  Per step: U/8 MMAs (M = 64, N = 128, kind::tf32) -> 16 epilogue warps (4 per TMEM sub-partition, rows
//   in lanes 0..15 of each: the M = 64 layout) apply the gates and store acts / c / h_t to global memory as the backward pass
//   needs them anyway -> the CTA then MULTICASTS its own chunk of h_t by TMA from global memory (its own stores, L2-resident)
//   into the A buffers of all 8 CTAs; the transfer completing on each CTA's mbarrier IS the signal that the chunk is there.
//   Write-after-read on the single A buffer: every CTA's MMAs of a step commit-multicast to the `a_free` barrier of all 8
//   CTAs; a CTA issues its multicast for step t only once all 8 have finished reading h_{t-1}.
//   16 clusters x 8 = 128 CTAs for cfg3 (B 1024, U 256); each SM ingests 64 KB per step instead of the 128 KB of the form above.
// =====================================================================================================================
constexpr int kLcThreads = 64 + 512;  // warp 0: TMA, warp 1: MMA issuer + TMEM owner, warps 2..17: epilogue
constexpr int kLcCluster = 8;
constexpr int kLcAChunk = 64 * 128;   // 64 rows x 32 k floats

__device__ __forceinline__ void lc_stamp(unsigned long long* dbg, int t, int slot) {
  if (dbg) {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    dbg[t * 8 + slot] = v;
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tma_load_2d_multicast(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
// arrives on `bar` (same offset) in every CTA of `mask` once this thread's earlier MMAs have completed
__device__ __forceinline__ void umma_commit_multicast_elect(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "h"(mask)
      : "memory");
}

__global__ void __launch_bounds__(kLcThreads, 1)
lstm_cluster_fwd_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_wr, const LstmPersistParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int kchunks = p.U >> 5;                               // = cluster size (the launcher checks U == 32 * kLcCluster)
  const int b_bytes = kchunks * 4 * 4096;                     // resident Wr slice: per 32-k chunk four [32 n x 32 k] boxes
  uint8_t* sm_b = smem;
  uint8_t* sm_a = smem + b_bytes;                             // [kchunks][64 rows x 128 B], 128-byte swizzled
  float* s_wk = reinterpret_cast<float*>(sm_a + kchunks * kLcAChunk);    // [F][128] input-kernel slice
  float* s_b = s_wk + kLpMaxF * 128;                                     // [128]   bias slice
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_b + 128);  // [2]: chunks of h_t land on a_full[t & 1]
  uint64_t* a_free = a_full + 2;                              // all CTAs of the cluster have finished reading the A buffer
  uint64_t* h_stored = a_free + 1;                            // this CTA's epilogue has stored (and fenced) its chunk of h_t
  uint64_t* tmem_full = h_stored + 1;
  uint64_t* tmem_empty = tmem_full + 1;
  uint64_t* b_bar = tmem_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int mi = blockIdx.x / kLcCluster;
  const int col0 = rank * 128, row0 = mi * 64;
  const uint16_t all = (uint16_t)((1u << kLcCluster) - 1);
  unsigned long long* dbg = blockIdx.x == 0 ? p.dbg : nullptr;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_h);
    tma_prefetch_desc(&map_wr);
    mbar_init(&a_full[0], 1);
    mbar_init(&a_full[1], 1);
    mbar_init(a_free, kLcCluster);
    mbar_init(h_stored, 512);
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 16);
    mbar_init(b_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();  // every CTA's barriers exist before any peer arrives on them or multicasts into this CTA
  pdl_wait();          // the pack kernels before this one wrote Wr_cat / Wk_cat / b_cat
  for (int i = threadIdx.x; i < p.F * 128; i += blockDim.x) s_wk[i] = p.wk_cat[(size_t)(i >> 7) * 4 * p.U + col0 + (i & 127)];
  if (threadIdx.x < 128) s_b[threadIdx.x] = p.has_bias ? p.b_cat[col0 + threadIdx.x] : 0.f;
  __syncthreads();
  const int t_first = p.zero_state ? 1 : 0;  // first step with a recurrent product

  if (warp == 0) {
    // ================= TMA: resident Wr slice, the initial state, then one multicast of this CTA's h chunk per step =================
    if (lane == 0) {
      mbar_arrive_expect_tx(b_bar, (uint32_t)b_bytes);
      for (int kc = 0; kc < kchunks; ++kc)
        for (int j = 0; j < 4; ++j) tma_load_2d(sm_b + (size_t)(kc * 4 + j) * 4096, &map_wr, b_bar, col0 + j * 32, kc * 32);
      if (t_first == 0) {  // stateful: h_{-1} (slot 0) into the A buffer, as if step -1 had produced it (barrier [1], like any odd step)
        mbar_arrive_expect_tx(&a_full[1], (uint32_t)(kchunks * kLcAChunk));
        for (int kc = 0; kc < kchunks; ++kc) tma_load_2d(sm_a + (size_t)kc * kLcAChunk, &map_h, &a_full[1], kc * 32, row0);
      }
      uint32_t hs_ph = 0, af_ph = 0;
      for (int t = 0; t + 1 < p.T; ++t) {   // the last step's h is not an operand of this kernel
        mbar_wait(h_stored, hs_ph);  hs_ph ^= 1;   // the 512 epilogue threads have stored + fenced their part of h_t
        lc_stamp(dbg, t, 6);
        mbar_wait(a_free, af_ph);    af_ph ^= 1;   // all 8 CTAs have finished the MMAs that read h_{t-1}
        asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy stores (global) -> this thread's async-proxy (TMA) loads
        mbar_arrive_expect_tx(&a_full[t & 1], (uint32_t)(kchunks * kLcAChunk));  // this CTA's barrier: all chunks of h_t
        tma_load_2d_multicast(sm_a + (size_t)rank * kLcAChunk, &map_h, &a_full[t & 1], rank * 32, (t + 1) * p.B + row0, all);
        lc_stamp(dbg, t, 7);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_tf32(64, 128, 0, 1);
    const uint64_t a_d = smem_desc_sw128(0, 16, 1024), b_d = smem_desc(0, 4096, 512, 1);
    const uint32_t a_hi = (uint32_t)(a_d >> 32), a_lo0 = (uint32_t)a_d, b_hi = (uint32_t)(b_d >> 32), b_lo0 = (uint32_t)b_d;
    // In a cluster launch a shared::cta address carries the CTA's rank in its upper bits; the descriptor's start-address field is
    // 14 bits of (offset >> 4) -- unmasked, the rank bits land in the LBO field (found as: gate columns 32.. of every CTA but rank 0
    // read their B groups from the wrong place)
    const uint32_t sb0 = b_lo0 + ((smem_u32(sm_b) >> 4) & 0x3FFFu), sa0 = a_lo0 + ((smem_u32(sm_a) >> 4) & 0x3FFFu);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    mbar_wait(b_bar, 0);
    tc_fence_after();
    uint32_t full_ph[2] = {0, 0}, eph = 0;
    for (int t = 0; t < p.T; ++t) {
      if (t >= t_first) {
        mbar_wait(tmem_empty, eph ^ 1);  // the epilogue has drained the previous step's accumulator
        eph ^= 1;
        const int fb = (t + 1) & 1;      // h_{t-1} was signalled on a_full[(t - 1) & 1]
        mbar_wait(&a_full[fb], full_ph[fb]);
        full_ph[fb] ^= 1;
        asm volatile("fence.proxy.async;" ::: "memory");
        tc_fence_after();
        if (lane == 0) lc_stamp(dbg, t, 0);
        uint32_t accumulate = 0;
        for (int kc = 0; kc < kchunks; ++kc) {
          const uint32_t at = sa0 + (uint32_t)(kc * (kLcAChunk >> 4));
          const uint32_t bt = sb0 + (uint32_t)(kc * (16384 >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_tf32_elect(tmem_u, ((uint64_t)a_hi << 32) | (at + ks * 2), ((uint64_t)b_hi << 32) | (bt + ks * (1024 >> 4)), idesc, accumulate);
            accumulate = 1;
          }
        }
        umma_commit_elect(tmem_full);
        if (lane == 0) lc_stamp(dbg, t, 1);
      }
      // "this CTA no longer reads the A buffer of step t" -> all 8 CTAs (immediately when the step had no recurrent product)
      umma_commit_multicast_elect(a_free, all);
    }
  } else {
    // ================= epilogue: 16 warps = TMEM sub-partition (warp % 4) x column quarter; rows in lanes 0..15 (M = 64) =================
    const int sub = warp & 3, quarter = (warp - 2) >> 2;
    const bool live = lane < 16;
    const int r = sub * 16 + (lane & 15), row = row0 + r;
    const int cbase = quarter * 32;           // this thread's 32 gate columns inside the CTA's 128 = 8 units
    const int u0 = (col0 + cbase) >> 2;       // first of its 8 units
    float cst[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cst[j] = (live && p.c_init) ? p.c_init[(size_t)row * p.U + u0 + j] : 0.f;
    uint32_t fph = 0;
    for (int t = 0; t < p.T; ++t) {
      float z[32];
      const bool has_z = t >= t_first;
      if (has_z) {
        mbar_wait(tmem_full, fph);
        fph ^= 1;
        tc_fence_after();
        if (threadIdx.x == 64) lc_stamp(dbg, t, 2);
        tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)cbase, z);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty);  // drained: the next step's MMAs may overwrite it
        if (threadIdx.x == 64) lc_stamp(dbg, t, 3);
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) z[q] = 0.f;
      }
      if (live) {
        float xv[kLpMaxF];
#pragma unroll
        for (int f = 0; f < kLpMaxF; ++f) xv[f] = f < p.F ? __ldg(p.x_tm + ((size_t)t * p.B + row) * p.F + f) : 0.f;
        float* arow = p.acts + ((size_t)t * p.B + row) * 4 * p.U + col0 + cbase;
        float hq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int cc = cbase + 4 * j + g;
            float xp = 0.f;
#pragma unroll
            for (int f = 0; f < kLpMaxF; ++f)
              if (f < p.F) xp = fmaf(xv[f], s_wk[f * 128 + cc], xp);
            float v = has_z ? __fadd_rn(xp, z[4 * j + g]) : xp;       // (x.Wk + h.Wr)
            if (p.has_bias) v = __fadd_rn(v, s_b[cc]);                  // + b
            a[g] = lp_act(v, g == 2 ? p.act : p.ract);
          }
          *reinterpret_cast<float4*>(arow + 4 * j) = make_float4(a[0], a[1], a[2], a[3]);
          const float cn = __fadd_rn(__fmul_rn(cst[j], a[0]), __fmul_rn(a[1], a[2]));   // c_prev*f + i*g
          cst[j] = cn;
          hq[j] = lp_tf32_rn(__fmul_rn(a[3], lp_act(cn, p.act)));                                    // o * act(c)
        }
        float* crow = p.c + ((size_t)t * p.B + row) * p.U + u0;
        float* hrow = p.h + ((size_t)(t + 1) * p.B + row) * p.U + u0;
        *reinterpret_cast<float4*>(hrow) = make_float4(hq[0], hq[1], hq[2], hq[3]);
        *reinterpret_cast<float4*>(hrow + 4) = make_float4(hq[4], hq[5], hq[6], hq[7]);
        *reinterpret_cast<float4*>(crow) = make_float4(cst[0], cst[1], cst[2], cst[3]);
        *reinterpret_cast<float4*>(crow + 4) = make_float4(cst[4], cst[5], cst[6], cst[7]);
      }
      // h_t of this thread is in global memory: make it visible to the TMA engine (async proxy) and tell the issuer
      if (threadIdx.x == 64) lc_stamp(dbg, t, 4);
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      if (threadIdx.x == 64) lc_stamp(dbg, t, 5);
      mbar_arrive(h_stored);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while a peer may still arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

struct LstmPersistPlan {
  CUtensorMap map_h, map_wr;
  LstmPersistParams p;
  int grid, smem;
  unsigned int* bar;
  int cluster = 0;  // 1: lstm_cluster_fwd_kernel (8-CTA clusters)
  unsigned long long* dbg = nullptr;
};

}  // namespace

// returns an opaque plan or null when the shape is not covered (the caller keeps the per-step path)
void* lstm_persist_create(int B, int T, int U, int F, float* h, const float* wr_cat, int device) {
  // OPT-IN (SB_LSTM_PERSIST=1).  Measured on cfg3 (B 1024, T 64, U 256; 128 CTAs): 12.1 us per step against 10.6 us for the per-step
  // GEMM + gate kernels (forward 775 vs 676 us) -- a step is a chain of global-memory round trips (h stores -> fence -> release add
  // -> acquire poll -> proxy fence -> TMA loads of the 128 KB row block) plus the gate math on only eight warps per SM; what would
  // pay is a cluster whose CTAs exchange h through distributed shared memory (profiles/r2_notes.md 11).
  const char* on = getenv("SB_LSTM_PERSIST");
  if (on && on[0] == '2') {  // cluster form
    if (B % 64 || U != 32 * kLcCluster || F > kLpMaxF || F < 1) return nullptr;
    LstmPersistPlan* P = new LstmPersistPlan();
    P->cluster = 1;
    uint64_t hd[2] = {(uint64_t)U, (uint64_t)(T + 1) * B};
    uint32_t hb[2] = {32, 64};
    P->map_h = make_tma_map(h, 2, hd, hb, false);
    uint64_t wd[2] = {(uint64_t)4 * U, (uint64_t)U};
    uint32_t wb[2] = {32, 32};
    P->map_wr = make_tma_map(wr_cat, 2, wd, wb, true);
    P->grid = (B / 64) * kLcCluster;
    P->smem = (U / 32) * 16384 + (U / 32) * kLcAChunk + (kLpMaxF * 128 + 128) * 4 + 16 * 16 + 2048;
    P->bar = nullptr;
    int dev_smem = 0;
    SB_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (P->smem > dev_smem) { delete P; return nullptr; }
    SB_CUDA(cudaFuncSetAttribute(lstm_cluster_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P->smem));
    // every cluster must be resident at once: the CTAs of a cluster wait for each other, and a cluster that is not scheduled yet
    // would only delay (not deadlock) the others -- but a grid that does not fit doubles the sequence time, so refuse it
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P->grid); cfg.blockDim = dim3(kLcThreads); cfg.dynamicSmemBytes = P->smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kLcCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int max_clusters = 0;
    const cudaError_t oc = cudaOccupancyMaxActiveClusters(&max_clusters, lstm_cluster_fwd_kernel, &cfg);
    { const char* v = getenv("SB_TC_DEBUG"); if (v && v[0] == '1') fprintf(stderr, "[sb] lstm cluster: grid %d CTAs, max active clusters %d (%s)\n", P->grid, max_clusters, cudaGetErrorString(oc)); }
    // the clusters are independent (row blocks): a grid of more clusters than fit runs in waves -- still correct, each wave costs a sequence
    if (oc != cudaSuccess || max_clusters < 1) {
      cudaGetLastError();
      delete P;
      return nullptr;
    }
    P->p.B = B; P->p.T = T; P->p.U = U; P->p.F = F; P->p.NB = kLcCluster;
    P->p.bar = nullptr;
    { const char* v = getenv("SB_LC_TIMELINE"); if (v && v[0] == '1') { SB_CUDA(cudaMallocManaged((void**)&P->dbg, (size_t)T * 8 * sizeof(unsigned long long))); memset(P->dbg, 0, (size_t)T * 8 * sizeof(unsigned long long)); } }
    P->p.dbg = P->dbg;
    return P;
  }
  if (!(on && on[0] == '1')) return nullptr;
  if (B % 128 || U % 32 || (4 * U) % 64 || U > 256 || F > kLpMaxF || F < 1) return nullptr;
  const int NB = 4 * U / 64, MB = B / 128;
  if (MB * NB > kNumSMs) return nullptr;  // every CTA must be resident: they wait for each other
  LstmPersistPlan* P = new LstmPersistPlan();
  uint64_t hd[2] = {(uint64_t)U, (uint64_t)(T + 1) * B};
  uint32_t hb[2] = {32, 128};
  P->map_h = make_tma_map(h, 2, hd, hb, false);
  uint64_t wd[2] = {(uint64_t)4 * U, (uint64_t)U};
  uint32_t wb[2] = {32, 32};
  P->map_wr = make_tma_map(wr_cat, 2, wd, wb, true);
  P->grid = MB * NB;
  P->smem = (U / 32) * 8192 + kLpStages * kLpATile + (kLpMaxF * 64 + 64) * 4 + 16 * 16 + 2048;
  int dev_smem = 0;
  SB_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  if (P->smem > dev_smem) { delete P; return nullptr; }
  SB_CUDA(cudaFuncSetAttribute(lstm_persist_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P->smem));
  SB_CUDA(cudaMalloc((void**)&P->bar, 64 * sizeof(unsigned int)));
  P->p.B = B; P->p.T = T; P->p.U = U; P->p.F = F; P->p.NB = NB;
  P->p.bar = P->bar;
  return P;
}

void lstm_persist_destroy(void* plan) {
  if (!plan) return;
  LstmPersistPlan* P = (LstmPersistPlan*)plan;
  if (P->bar) cudaFree(P->bar);
  if (P->dbg) {  // developer timeline: mean interval between the stamps of CTA 0 over the last launch
    cudaDeviceSynchronize();
    const int T = P->p.T;
    const char* nm[8] = {"mma_start", "mma_issued", "tmem_full_seen", "tmem_ld_done", "gates+stores_done", "fences_done", "h_stored_seen", "multicast_issued"};
    double acc[8] = {0}; int cnt = 0;
    for (int t = 2; t + 2 < T; ++t) {
      const unsigned long long* d = P->dbg + (size_t)t * 8;
      if (!d[0] || !d[7]) continue;
      for (int k = 0; k < 8; ++k) acc[k] += (double)(long long)(d[k] - d[0]);
      ++cnt;
    }
    fprintf(stderr, "[sb] lstm cluster timeline (ns after the step's mma_start, mean of %d steps):", cnt);
    for (int k = 0; k < 8; ++k) fprintf(stderr, " %s %.0f", nm[k], cnt ? acc[k] / cnt : 0.0);
    if (T > 6) fprintf(stderr, " | step-to-step %.0f\n", (double)(long long)(P->dbg[(size_t)(T - 3) * 8] - P->dbg[(size_t)2 * 8]) / (T - 5));
    cudaFree(P->dbg);
  }
  delete P;
}

void lstm_persist_forward(void* plan, const float* x_tm, const float* wk_cat, const float* b_cat, const float* c_init, float* acts,
                          float* c, float* h, int act, int ract, int has_bias, int zero_state, cudaStream_t s) {
  LstmPersistPlan* P = (LstmPersistPlan*)plan;
  LstmPersistParams p = P->p;
  p.x_tm = x_tm; p.wk_cat = wk_cat; p.b_cat = b_cat; p.c_init = c_init; p.acts = acts; p.c = c; p.h = h;
  p.act = act; p.ract = ract; p.has_bias = has_bias; p.zero_state = zero_state;
  if (P->cluster) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P->grid); cfg.blockDim = dim3(kLcThreads); cfg.dynamicSmemBytes = P->smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kLcCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 2;
    SB_CUDA(cudaLaunchKernelEx(&cfg, lstm_cluster_fwd_kernel, P->map_h, P->map_wr, p));
    set_prewait_ok(1);
    SB_LAUNCHED();
    return;
  }
  SB_CUDA(cudaMemsetAsync(P->bar, 0, 64 * sizeof(unsigned int), s));
  launch_k(lstm_persist_fwd_kernel, P->grid, kLpThreads, P->smem, s, P->map_h, P->map_wr, p);
  SB_LAUNCHED();
}

void trace_set_lstm_persist(unsigned long long* p) { sb_trace_set_local(p); }  // SB_TRACE timeline buffer of this translation unit

}  // namespace sb
