// tcgen05 / TMEM / TMA dense GEMM of the scanet-b200 engine (sm_100a), kind::tf32 with fp32 accumulation:
//   C[M,N] = op(A) . op(B)      -- MatMul forward and its two gradients (math/alg/AllKernels.scala:370-389)
// One persistent kernel serves the three operand layouts of the train step
//   NN  Z  = X . W        A K-major  [M][K],  B MN-major [K][N]     (Dense forward, LSTM gates)
//   NT  dX = G . W^T      A K-major  [M][K],  B K-major  [N][K]     (Dense dgrad, LSTM dH)
//   TN  dW = X^T . G      A MN-major [K][M],  B MN-major [K][N]     (Dense wgrad, LSTM dW over the whole sequence)
// Tile 128 x BN x 32 (BN <= 256).  Warp 0 = TMA producer, warp 1 = MMA issuer (warp-uniform loop, elected lane), warps 2..5 =
// epilogue: TMEM -> registers -> (+bias, activation) -> 128B-swizzled shared staging -> TMA store (edge tiles are clipped by
// the tensor map, K tails are zero-filled by the loads).  Accumulators are double-buffered in TMEM so the epilogue of tile i
// overlaps the MMAs of tile i+1.  Optional deterministic split-K: partial tiles go to slices of a workspace tensor and are
// summed in split order by partial_reduce.  All tensor maps are 3-D; the outer coordinate selects a slice (an LSTM time step,
// a split-K partial), so one plan serves every step of a recurrence.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <stdexcept>
#include <string>

#include "common.cuh"
#include "kernels.cuh"
#include "tc.h"
#include "tc_sm100.cuh"

namespace sb {

using namespace tc;

struct GemmParams {
  int M, N, K;
  int BN;
  int a_mn_major, b_mn_major;
  int a_grouped, b_grouped;  // MN-major operand fetched as ONE 4-D box [32, 32 k, groups, 1] (inner dim % 32 == 0)
  int m_tiles, n_tiles, splits, kb_total, kb_per_split;
  int stages;
  int za, zb, zc;
  int x3;   // 3xTF32: split operands (hi = tf32(x), lo = x - hi), D += A_hi.B_lo + A_lo.B_hi + A_hi.B_hi  (fp32-class accuracy)
  int act;  // sb_activation applied in the epilogue (0 none, 1 sigmoid, 2 tanh, 4 relu)
  float thr;
  const float* bias;  // [N] or null
  LstmEpi le;         // le.mode != 0: LSTM gate math in the epilogue (tc.h)
};

constexpr int kGemmThreads = 192;       // producer, MMA issuer, 4 epilogue warps
constexpr int kGemmThreadsX3 = 320;     // + 4 converter warps that materialise the low-order operand tiles
constexpr int kGemmATile = 128 * 128;  // 128 rows x 32 fp32

__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_g(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float epi_act(float v, int act, float thr) {
  switch (act) {
    case 1: return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-v)));
    case 2: return tanhf(v);
    case 4: return fmaxf(thr, v);
    default: return v;
  }
}

// Tile order: groups of kGroupM row tiles sweep all column tiles before the next group starts, so the CTAs running at any
// moment share a small set of A row panels (L2-resident) instead of streaming all of A once per column tile.
constexpr int kGroupM = 16;
__device__ __forceinline__ void tile_to_mn(int tile, int m_tiles, int n_tiles, int& mt, int& nt) {
  const int per_group = kGroupM * n_tiles;
  const int grp = tile / per_group, in = tile - grp * per_group;
  const int gm = min(kGroupM, m_tiles - grp * kGroupM);  // rows in this (possibly last, shorter) group
  nt = in / gm;
  mt = grp * kGroupM + (in - nt * gm);
}

// ---- LSTM epilogues (layer/RNN.scala:319-339; the same arithmetic as lstm_gates_fwd/bwd_kernel in lstm.cu) ----
__device__ __forceinline__ float lstm_actf(float x, int act) {
  switch (act) {
    case 1: return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
    case 2: return tanhf(x);
    case 4: return fmaxf(0.f, x);
    default: return x;
  }
}
__device__ __forceinline__ float lstm_actg(float g, float y, int act) {
  switch (act) {
    case 1: return __fmul_rn(__fmul_rn(y, __fsub_rn(1.f, y)), g);
    case 2: return __fmul_rn(__fsub_rn(1.f, __fmul_rn(y, y)), g);
    case 4: return y > 0.f ? g : 0.f;
    default: return g;
  }
}
// TF32 mode: the approximate forms of lstm.cu actf_t (ex2.approx / rcp.approx)
__device__ __forceinline__ float lstm_actf_fast(float x, int act) {
  switch (act) {
    case 1: return __fdividef(1.f, 1.f + __expf(-x));
    case 2: return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x));
    case 4: return fmaxf(0.f, x);
    default: return x;
  }
}
// v[32] = z of row b, columns col0..col0+31 = units col0/4 .. col0/4+7 (4 gates each): becomes the activations
__device__ __forceinline__ void lstm_epi_fwd(float* v, long long b, int col0, const LstmEpi& le) {
  const int j0 = col0 >> 2;
  float cp[8], cn[8], hn[8];
  *reinterpret_cast<float4*>(cp) = *reinterpret_cast<const float4*>(le.c_prev + b * le.U + j0);
  *reinterpret_cast<float4*>(cp + 4) = *reinterpret_cast<const float4*>(le.c_prev + b * le.U + j0 + 4);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    float a[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = col0 + 4 * u + g;
      float xp = 0.f;
      for (int f = 0; f < le.F; ++f) xp = fmaf(__ldg(le.x_t + b * le.F + f), __ldg(le.wk_cat + (long long)f * 4 * le.U + col), xp);
      float z = __fadd_rn(xp, v[4 * u + g]);                    // (x.Wk + h.Wr)
      if (le.has_bias) z = __fadd_rn(z, __ldg(le.b_cat + col));  // + b
      a[g] = le.fast ? lstm_actf_fast(z, g == 2 ? le.act : le.ract) : lstm_actf(z, g == 2 ? le.act : le.ract);
      v[4 * u + g] = a[g];
    }
    cn[u] = __fadd_rn(__fmul_rn(cp[u], a[0]), __fmul_rn(a[1], a[2]));  // c_prev*f + i*g
    hn[u] = __fmul_rn(a[3], le.fast ? lstm_actf_fast(cn[u], le.act) : lstm_actf(cn[u], le.act));  // o * act(c)
    if (le.fast) { unsigned r_; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r_) : "f"(hn[u])); hn[u] = __uint_as_float(r_); }  // lstm.cu tf32_rn
  }
  *reinterpret_cast<float4*>(le.c_out + b * le.U + j0) = *reinterpret_cast<float4*>(cn);
  *reinterpret_cast<float4*>(le.c_out + b * le.U + j0 + 4) = *reinterpret_cast<float4*>(cn + 4);
  *reinterpret_cast<float4*>(le.h_out + b * le.U + j0) = *reinterpret_cast<float4*>(hn);
  *reinterpret_cast<float4*>(le.h_out + b * le.U + j0 + 4) = *reinterpret_cast<float4*>(hn + 4);
}
// v[32] = dH_{t-1} of row b, units j0..j0+31: gate backward of step t-1 (oracle/layers.py LSTMCell.step_bwd)
__device__ __forceinline__ void lstm_epi_bwd(const float* v, long long b, int j0, const LstmEpi& le) {
#pragma unroll 4
  for (int u = 0; u < 32; ++u) {
    const long long i = b * le.U + j0 + u;
    float4* q = reinterpret_cast<float4*>(le.acts_dz + b * 4 * le.U + 4 * (j0 + u));
    const float4 a4 = *q;
    const float f = a4.x, ig = a4.y, g = a4.z, o = a4.w;
    const float ac = le.fast ? lstm_actf_fast(le.c_cur[i], le.act) : lstm_actf(le.c_cur[i], le.act);
    const float dhv = v[u];
    const float d_o = __fmul_rn(dhv, ac);
    const float dc = __fadd_rn(lstm_actg(__fmul_rn(dhv, o), ac, le.act), le.dc_io[i]);
    const float cp = le.c_prev ? le.c_prev[i] : 0.f;
    *q = make_float4(lstm_actg(__fmul_rn(dc, cp), f, le.ract), lstm_actg(__fmul_rn(dc, g), ig, le.ract),
                     lstm_actg(__fmul_rn(dc, ig), g, le.act), lstm_actg(d_o, o, le.ract));
    le.dc_io[i] = __fmul_rn(dc, f);
  }
}

__global__ void __launch_bounds__(kGemmThreadsX3, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_tile = p.BN * 128;
  const int hi_bytes = kGemmATile + b_tile;                  // [A | B] as loaded by TMA
  const int stage_bytes = p.x3 ? 2 * hi_bytes : hi_bytes;    // 3xTF32: followed by [A_lo | B_lo]
  uint8_t* sm_epi = smem + (size_t)p.stages * stage_bytes;  // 4 warps x 2 buffers x 4 KB
  uint64_t* full_bar = (uint64_t*)(sm_epi + 4 * 2 * 4096);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* conv_bar = empty_bar + p.stages;  // 3xTF32: low-order tiles of the stage are written
  uint64_t* tmem_full = conv_bar + p.stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // shfl: provably warp-uniform
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * p.BN) tmem_cols <<= 1;
  const int tiles = p.m_tiles * p.n_tiles;
  const int n_work = tiles * p.splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_c);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&conv_bar[s], 4);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the on-chip setup above overlapped the previous kernel; global memory from here on
  // (no early launch trigger here: measured on cfg3 / cfg4, waiting dependents cost this persistent GEMM 2 %; profiles/r1_notes.md 17)

  if (warp >= 6) {
    // ================= 3xTF32 converters: lo = x - tf32(x) for every word of the stage's A and B tiles =================
    // kind::tf32 reads fp32 words and ignores the low 13 mantissa bits, so the hi tiles ARE the tiles TMA wrote; the lo
    // tiles have the same (swizzled) layout, word for word.
    const int t = threadIdx.x - 192;
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int split = w / tiles;
      const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        const float4* hi = reinterpret_cast<const float4*>(smem + (size_t)stage * stage_bytes);
        float4* lo = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes + hi_bytes);
        for (int i = t; i < hi_bytes / 16; i += 128) {
          const float4 v = hi[i];
          float4 r;
          r.x = __fsub_rn(v.x, __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
          r.y = __fsub_rn(v.y, __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
          r.z = __fsub_rn(v.z, __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
          r.w = __fsub_rn(v.w, __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
          lo[i] = r;
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int split = w / tiles, tile = w - split * tiles;
        int mt, nt;
        tile_to_mn(tile, p.m_tiles, p.n_tiles, mt, nt);
        const int m0 = mt * 128, n0 = nt * p.BN;
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          uint8_t* sb = sa + kGemmATile;
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)hi_bytes);
          if (p.a_grouped) {
            tma_load_4d_g(sa, &map_a, &full_bar[stage], 0, kb * 32, m0 >> 5, p.za);
          } else if (p.a_mn_major) {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_3d(sa + j * 4096, &map_a, &full_bar[stage], m0 + j * 32, kb * 32, p.za);
          } else {
            tma_load_3d(sa, &map_a, &full_bar[stage], kb * 32, m0, p.za);
          }
          if (p.b_grouped) {
            tma_load_4d_g(sb, &map_b, &full_bar[stage], 0, kb * 32, n0 >> 5, p.zb);
          } else if (p.b_mn_major) {
            for (int j = 0; j < p.BN / 32; ++j) tma_load_3d(sb + j * 4096, &map_b, &full_bar[stage], n0 + j * 32, kb * 32, p.zb);
          } else {
            tma_load_3d(sb, &map_b, &full_bar[stage], kb * 32, n0, p.zb);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: warp-uniform loop, one elected lane issues =================
    const uint32_t idesc = idesc_tf32(128, p.BN, p.a_mn_major, p.b_mn_major);
    const uint64_t a_d = p.a_mn_major ? smem_desc(0, 4096, 512, 1) : smem_desc_sw128(0, 16, 1024);
    const uint64_t b_d = p.b_mn_major ? smem_desc(0, 4096, 512, 1) : smem_desc_sw128(0, 16, 1024);
    const uint32_t a_hi = (uint32_t)(a_d >> 32), a_lo0 = (uint32_t)a_d, b_hi = (uint32_t)(b_d >> 32), b_lo0 = (uint32_t)b_d;
    const uint32_t a_step = p.a_mn_major ? (1024u >> 4) : (32u >> 4), b_step = p.b_mn_major ? (1024u >> 4) : (32u >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int split = w / tiles;
      const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
      uint32_t accumulate = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(p.x3 ? &conv_bar[stage] : &full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes) >> 4;
        const uint32_t at = a_lo0 + sa, bt = b_lo0 + sa + (uint32_t)(kGemmATile >> 4);
        if (p.x3) {
          const uint32_t lo_off = (uint32_t)(hi_bytes >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // small terms first: A_hi.B_lo, A_lo.B_hi, then A_hi.B_hi
            const uint64_t da = ((uint64_t)a_hi << 32) | (at + ks * a_step), db = ((uint64_t)b_hi << 32) | (bt + ks * b_step);
            umma_tf32_elect(d_tmem, da, db + lo_off, idesc, accumulate);
            umma_tf32_elect(d_tmem, da + lo_off, db, idesc, 1u);
            umma_tf32_elect(d_tmem, da, db, idesc, 1u);
            accumulate = 1;
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // 4 x (K = 8 tf32) per 32-wide k block
            umma_tf32_elect(d_tmem, ((uint64_t)a_hi << 32) | (at + ks * a_step), ((uint64_t)b_hi << 32) | (bt + ks * b_step), idesc,
                            accumulate);
            accumulate = 1;
          }
        }
        umma_commit_elect(&empty_bar[stage]);  // the smem slot is free once these MMAs have read it
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit_elect(&tmem_full[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================= epilogue: TMEM -> registers -> (+bias, act) -> swizzled smem -> TMA store =================
    const int sub = warp & 3;  // TMEM sub-partition this warp may read: lanes [32*sub, 32*sub+32)
    uint8_t* stg = sm_epi + (size_t)(warp - 2) * 2 * 4096;
    int buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int split = w / tiles, tile = w - split * tiles;
      int mt, nt;
      tile_to_mn(tile, p.m_tiles, p.n_tiles, mt, nt);
      const int m0 = mt * 128 + sub * 32, n0 = nt * p.BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * p.BN + c0), v);
        tmem_ld_wait();
        if (p.le.mode == 2) {  // LSTM backward: dH is consumed here, nothing is stored
          if (n0 + c0 < p.N && m0 + lane < p.M) lstm_epi_bwd(v, m0 + lane, n0 + c0, p.le);
          continue;
        }
        if (n0 + c0 < p.N) {  // whole 32-column chunks beyond N are skipped (the store would be clipped anyway)
          if (p.le.mode == 1 && m0 + lane < p.M) lstm_epi_fwd(v, m0 + lane, n0 + c0, p.le);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __fadd_rn(v[j], (n0 + c0 + j < p.N) ? __ldg(p.bias + n0 + c0 + j) : 0.f);
          }
          if (p.act) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = epi_act(v[j], p.act, p.thr);
          }
          // the staging buffer is free once the store issued two chunks ago has READ it
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* sbuf = stg + buf * 4096;
          // row = lane (128 B), 16-byte chunk j at position j ^ (lane & 7): the 128B swizzle the store's tensor map expects
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(sbuf + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_c, sbuf, n0 + c0, m0, p.zc + split);
            tma_store_commit();
          }
          buf ^= 1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();  // global writes complete before the kernel ends
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// =====================================================================================================================
// host side
// =====================================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    SB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled is not available");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}
// fp32 [slices][outer][inner] tensor, box {32, box_outer, 1}
static CUtensorMap map3(const float* base, uint64_t inner, uint64_t outer, uint64_t slices, uint32_t box_outer, bool atom32) {
  CUtensorMap m;
  cuuint64_t gdim[3] = {inner, outer, slices};
  cuuint64_t gstr[2] = {inner * 4, inner * outer * 4};
  cuuint32_t box[3] = {32, box_outer, 1}, es[3] = {1, 1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

// MN-major operand [slices][K][inner] with inner % 32 == 0, viewed as [slices][inner/32][K][32]: one box {32, 32 k, groups, 1}
// lands in shared memory as `groups` consecutive [32 k x 128 B] blocks -- exactly the UMMA MN-major tile (LBO = 4096)
static CUtensorMap map4_grouped(const float* base, uint64_t inner, uint64_t K, uint64_t slices, uint32_t groups) {
  CUtensorMap m;
  cuuint64_t gdim[4] = {32, K, inner / 32, slices};
  cuuint64_t gstr[3] = {inner * 4, 128, inner * K * 4};
  cuuint32_t box[4] = {32, 32, groups, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled (grouped) failed with code " + std::to_string((int)r));
  return m;
}

struct TcGemm {
  CUtensorMap map_a, map_b, map_c;
  GemmParams p{};
  int grid = 0, smem = 0;
  float* partial = nullptr;  // split-K workspace (borrowed)
  float* c = nullptr;
  long long c_slice = 0;
};

TcGemm* tc_gemm_create(int a_mn_major, int b_mn_major, const float* A, int a_slices, const float* B, int b_slices, float* C,
                       int c_slices, int M, int N, int K, int allow_split, float* ws, size_t ws_floats, int device, int x3) {
  // 16-byte global strides / base addresses
  const int a_inner = a_mn_major ? M : K, b_inner = b_mn_major ? N : K;
  if ((a_inner & 3) || (b_inner & 3) || (N & 3) || M < 1 || N < 8 || K < 1) return nullptr;
  if ((((uintptr_t)A) | ((uintptr_t)B) | ((uintptr_t)C)) & 15) return nullptr;
  int dev_smem = 0;
  SB_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  TcGemm* g = new TcGemm();
  GemmParams& p = g->p;
  p.M = M; p.N = N; p.K = K; p.a_mn_major = a_mn_major; p.b_mn_major = b_mn_major; p.x3 = x3 ? 1 : 0;
  p.m_tiles = (M + 127) / 128;
  p.kb_total = (K + 31) / 32;
  // tile width / split-K: maximise (SM utilisation over whole waves) x (MMA efficiency of the tile width)
  const int n_cap = (N + 31) / 32 * 32;
  double best = -1.0;
  for (int bn : {256, 128, 64, 32}) {
    if (x3 && bn > 128) continue;  // the low-order tiles double the stage: keep >= 3 stages
    const int BNc = std::min(bn, n_cap);
    const int tiles = p.m_tiles * ((N + BNc - 1) / BNc);
    int splits = 1;
    // allow_split == 2: a latency-critical small-output product on the serial path of a recurrence (LSTM dH = dZ . Wr^T: 16 tiles
    // of K = 1024) -- split down to 4 k-blocks per CTA so that every SM streams a share of the operands
    const int min_kb = allow_split == 2 ? 4 : 16;
    if (allow_split && ws && tiles * 2 <= kNumSMs && p.kb_total >= 4 * min_kb)
      splits = std::max(1, std::min({kNumSMs / tiles, p.kb_total / min_kb, (int)(ws_floats / ((size_t)M * N))}));
    const int work = tiles * splits, waves = (work + kNumSMs - 1) / kNumSMs;
    const double eff = BNc >= 256 ? 1.0 : BNc >= 128 ? 0.9 : BNc >= 64 ? 0.7 : 0.45;
    const double score = (double)work / (waves * (double)kNumSMs) * eff;
    if (score > best) { best = score; p.BN = BNc; p.splits = splits; }
  }
  const char* force = getenv("SB_TC_GEMM_BN");
  if (force && atoi(force) >= 32) { p.BN = std::min(atoi(force), 256); p.splits = 1; }
  p.n_tiles = (N + p.BN - 1) / p.BN;
  const int tiles = p.m_tiles * p.n_tiles;
  p.kb_per_split = (p.kb_total + p.splits - 1) / p.splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  const int stage_bytes = (x3 ? 2 : 1) * (kGemmATile + p.BN * 128);
  const int fixed = 4 * 2 * 4096 + 16 * 24 + 64 + 1024;
  p.stages = std::min(8, (dev_smem - fixed) / stage_bytes);
  if (p.stages < 2) { delete g; return nullptr; }
  g->smem = fixed + p.stages * stage_bytes;
  g->grid = std::min(tiles * p.splits, kNumSMs);
  const char* ng = getenv("SB_TC_GEMM_NO_GROUPED");
  const bool grouped_ok = !(ng && ng[0] == '1');
  p.a_grouped = a_mn_major && grouped_ok && (M % 32 == 0);
  p.b_grouped = b_mn_major && grouped_ok && (N % 32 == 0);
  g->map_a = p.a_grouped ? map4_grouped(A, M, K, a_slices, 4)
                         : (a_mn_major ? map3(A, M, K, a_slices, 32, true) : map3(A, K, M, a_slices, 128, false));
  g->map_b = p.b_grouped ? map4_grouped(B, N, K, b_slices, (uint32_t)(p.BN / 32))
                         : (b_mn_major ? map3(B, N, K, b_slices, 32, true) : map3(B, K, N, b_slices, (uint32_t)p.BN, false));
  g->c = C;
  g->c_slice = (long long)M * N;
  if (p.splits > 1) {
    g->partial = ws;
    g->map_c = map3(ws, N, M, p.splits, 32, false);
  } else {
    g->map_c = map3(C, N, M, c_slices, 32, false);
  }
  static bool attr = false;
  if (!attr) {
    SB_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    attr = true;
  }
  return g;
}

void tc_gemm_set_epilogue(TcGemm* g, const float* bias, int act, float thr) {
  if (g->p.splits > 1 && (bias || act)) throw std::runtime_error("fused epilogue is not available with split-K");
  g->p.bias = bias; g->p.act = act; g->p.thr = thr;
}
bool tc_gemm_is_split(const TcGemm* g) { return g->p.splits > 1; }

void tc_gemm_run(TcGemm* g, int za, int zb, int zc, cudaStream_t s) {
  GemmParams p = g->p;
  p.za = za; p.zb = zb; p.zc = p.splits > 1 ? 0 : zc;
  launch_k(gemm_tf32_kernel, g->grid, p.x3 ? kGemmThreadsX3 : kGemmThreads, g->smem, s, g->map_a, g->map_b, g->map_c, p);
  SB_LAUNCHED();
  if (p.splits > 1) partial_reduce(g->partial, g->c + (long long)zc * g->c_slice, p.M * p.N, p.splits, s);
}

// split-K plans only: the partial products stay in the workspace ([splits][M*N], to be summed in split order by the consumer)
int tc_gemm_run_partials(TcGemm* g, int za, int zb, cudaStream_t s, const float** partial) {
  GemmParams p = g->p;
  p.za = za; p.zb = zb; p.zc = 0;
  launch_k(gemm_tf32_kernel, g->grid, p.x3 ? kGemmThreadsX3 : kGemmThreads, g->smem, s, g->map_a, g->map_b, g->map_c, p);
  SB_LAUNCHED();
  *partial = g->partial;
  return p.splits;
}

void tc_gemm_run_lstm(TcGemm* g, int za, int zb, int zc, const LstmEpi& epi, cudaStream_t s) {
  if (g->p.splits > 1) throw std::runtime_error("LSTM epilogue is not available with split-K");
  if (epi.mode == 1 && (g->p.N != 4 * epi.U || (epi.U & 7))) throw std::runtime_error("LSTM forward epilogue: N must be 4U, U % 8 == 0");
  if (epi.mode == 2 && (g->p.N != epi.U || (epi.U & 31))) throw std::runtime_error("LSTM backward epilogue: N must be U, U % 32 == 0");
  GemmParams p = g->p;
  p.za = za; p.zb = zb; p.zc = zc;
  p.le = epi;
  launch_k(gemm_tf32_kernel, g->grid, p.x3 ? kGemmThreadsX3 : kGemmThreads, g->smem, s, g->map_a, g->map_b, g->map_c, p);
  SB_LAUNCHED();
}

void tc_gemm_destroy(TcGemm* g) { delete g; }

}  // namespace sb

namespace sb {
void trace_set_tc_gemm(unsigned long long* p) { sb_trace_set_local(p); }  // SB_TRACE timeline buffer of this translation unit
}  // namespace sb
