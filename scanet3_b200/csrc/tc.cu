// tcgen05 / TMEM / TMA kernels of the scanet-b200 engine (sm_100a): the true dense contractions of the train step.
//
//  * conv_igemm_kernel  -- Conv2D fprop and dgrad (stride 1, VALID) as implicit GEMMs (math/nn/AllKernels.scala:143-261):
//      an M-tile is BH full rows of one image's output grid (<= 128 pixels); for every filter tap the A operand is ONE
//      4-D TMA box over the NHWC source shifted by the tap (out-of-bounds pixels are zero-filled by TMA, which is exactly
//      the zero padding dgrad needs), the B operand is a slice of the HWIO filter; fp32 accumulators live in TMEM
//      (double-buffered) and a 4-warp epilogue applies ReLU and stores NHWC rows.
//  * conv_wgrad_kernel  -- Conv2DBackpropFilter: D[Cout, Cin] per tap, K = pixels; both operands MN-major straight out
//      of NHWC memory; per-CTA partial sums are reduced in a fixed order (deterministic).
// Arithmetic: kind::tf32 (operands are fp32 words, the low 13 mantissa bits are ignored), fp32 accumulation.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"
#include "engine.h"
#include "tc.h"
#include "tc_sm100.cuh"

namespace sb {

using namespace tc;

// =====================================================================================================================
// TMA descriptors
// =====================================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (!p || qres != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled is not available");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}
// fp32 tensor, dims innermost first, 128-byte swizzle, zero fill for out-of-bounds elements
static CUtensorMap make_map(const float* base, int rank, const uint64_t* dims, const uint32_t* box, bool mn_major = false) {
  CUtensorMap m;
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t bdim[5], estr[5];
  uint64_t stride = sizeof(float);
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    stride *= dims[i];
    if (i < rank - 1) gstride[i] = stride;
  }
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, gdim, gstride, bdim, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

// the same helper for the other translation units (lstm_persist.cu)
CUtensorMap make_tma_map(const float* base, int rank, const uint64_t* dims, const uint32_t* box, bool mn_major) {
  return make_map(base, rank, dims, box, mn_major);
}

// =====================================================================================================================
// implicit-GEMM Conv2D fprop / dgrad
// =====================================================================================================================
struct IgemmParams {
  int n_tiles, tiles_per_img;
  int Hg, Wg, BH;     // output grid of the tile space; tile = BH grid rows (BH*Wg <= 128 pixels)
  int KH, KW;
  int kchunks;        // 32-channel K chunks per tap
  int sign;           // +1 fprop: source pixel = (h+kh, w+kw) ; -1 dgrad: (h-kh, w-kw)
  int N;              // accumulator columns (fprop: Cout, dgrad: Cin), multiple of 32
  int b_mn_major;     // fprop 1: B = N/32 boxes [32 n x 32 k] ; dgrad 0: B = one box [32 k x N rows]
  int b_rows_per_tap; // row offset between taps in the 2-D filter view
  int stages;
  int relu;
  float thr;
  float* out;         // [n_img*Hg*Wg, ldc]: this launch writes columns [n0, n0 + N)
  int ldc, n0;        // output row pitch (all channels) and the first output channel of this launch
  int pt, pl;         // padding before (top, left) of the source grid: taps read (h + sign*kh - pt, w + sign*kw - pl)
  int x3;             // 3xTF32: converter warps write lo = x - tf32(x) tiles behind the stage's [A | B]; D += A_hi.B_lo + A_lo.B_hi + A_hi.B_hi
  // Small grids (Hg*Wg <= 64): a tile is `ipt` whole images (BH == Hg; the TMA box spans ipt images), so the M = 128 rows of the MMA
  // are filled instead of holding one 4x4 or 6x6 image.  nsplit > 1 (fprop): work unit = (tile, slice of N output channels), N is the
  // slice width -- more CTAs busy when there are few tiles, and every CTA streams only its slice of the filter.
  int ipt, nsplit, n_img;
};

constexpr int kIgemmThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..5: epilogue
constexpr int kIgemmThreadsX3 = 320;  // + warps 6..9: 3xTF32 converters (lo = x - tf32(x) for every word of a stage)
// lo = x - tf32(x): kind::tf32 ignores the low 13 mantissa bits, so the tile TMA wrote IS the hi operand; the lo tile has the same
// (swizzled) layout word for word.  128 threads walk `bytes` (multiple of 16).
__device__ __forceinline__ void split_lo_tile(const uint8_t* hi_p, uint8_t* lo_p, int bytes, int t) {
  const float4* hi = reinterpret_cast<const float4*>(hi_p);
  float4* lo = reinterpret_cast<float4*>(lo_p);
  for (int i = t; i < bytes / 16; i += 128) {
    const float4 v = hi[i];
    float4 r;
    r.x = __fsub_rn(v.x, __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
    r.y = __fsub_rn(v.y, __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
    r.z = __fsub_rn(v.z, __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
    r.w = __fsub_rn(v.w, __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
    lo[i] = r;
  }
}
constexpr int kATileBytes = 128 * 128;

__global__ void __launch_bounds__(kIgemmThreadsX3, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const IgemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A 16 KB | B N*128 B  (| A_lo | B_lo  in 3xTF32 mode))] then barriers
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_tile_bytes = p.N * 128;
  const int hi_bytes = kATileBytes + b_tile_bytes;
  const int stage_bytes = p.x3 ? 2 * hi_bytes : hi_bytes;
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* conv_bar = empty_bar + p.stages;  // 3xTF32: the stage's low-order tiles are written
  uint64_t* tmem_full = conv_bar + p.stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // shfl: provably warp-uniform
  const int a_box_bytes = p.BH * p.Wg * 128 * p.ipt;
  const int kblocks = p.KH * p.KW * p.kchunks;
  const int n_units = p.n_tiles * p.nsplit;  // unit = tile * nsplit + slice: the slices of a tile run back to back (A stays in L2)
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * p.N) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
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
  pdl_trigger();  // dependents may start their own on-chip setup now (they still wait for this grid before touching its output)

  if (warp >= 6) {
    // ================= 3xTF32 converters (launched only in that mode) =================
    const int t = threadIdx.x - 192;
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x)
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        uint8_t* st = smem + (size_t)stage * stage_bytes;
        split_lo_tile(st, st + hi_bytes, hi_bytes, t);
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
  } else if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int tile = unit / p.nsplit, n0 = p.n0 + (unit - tile * p.nsplit) * p.N;
        const int img = p.ipt > 1 ? tile * p.ipt : tile / p.tiles_per_img, h0 = p.ipt > 1 ? 0 : (tile % p.tiles_per_img) * p.BH;
        int tap = 0;
        for (int kh = 0; kh < p.KH; ++kh)
          for (int kw = 0; kw < p.KW; ++kw, ++tap)
            for (int kc = 0; kc < p.kchunks; ++kc) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + (size_t)stage * stage_bytes;
              uint8_t* sb = sa + kATileBytes;
              mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(a_box_bytes + b_tile_bytes));
              tma_load_4d(sa, &map_a, &full_bar[stage], kc * 32, p.sign * kw - p.pl, h0 + p.sign * kh - p.pt, img);
              if (p.b_mn_major) {
                for (int j = 0; j < p.N / 32; ++j)
                  tma_load_2d(sb + j * 4096, &map_b, &full_bar[stage], n0 + j * 32, tap * p.b_rows_per_tap + kc * 32);
              } else {
                tma_load_2d(sb, &map_b, &full_bar[stage], kc * 32, tap * p.b_rows_per_tap + n0);
              }
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: warp-uniform loop, one elected lane issues =================
    {
      const uint32_t idesc = idesc_tf32(128, p.N, 0, p.b_mn_major);
      const uint64_t a_d = smem_desc_sw128(0, 16, 1024), b_d = p.b_mn_major ? smem_desc(0, 4096, 512, 1) : smem_desc_sw128(0, 16, 1024);
      const uint32_t a_hi = (uint32_t)(a_d >> 32), a_lo0 = (uint32_t)a_d, b_hi = (uint32_t)(b_d >> 32), b_lo0 = (uint32_t)b_d;
      const uint32_t b_ks_step = p.b_mn_major ? (1024u >> 4) : (32u >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.N);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(p.x3 ? &conv_bar[stage] : &full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes) >> 4;
          const uint32_t at = a_lo0 + sa, bt = b_lo0 + sa + (uint32_t)(kATileBytes >> 4);
          if (p.x3) {
            const uint32_t lo_off = (uint32_t)(hi_bytes >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {  // small terms first: A_hi.B_lo, A_lo.B_hi, then A_hi.B_hi
              const uint64_t da = ((uint64_t)a_hi << 32) | (at + ks * 2), db = ((uint64_t)b_hi << 32) | (bt + ks * b_ks_step);
              umma_tf32_elect(d_tmem, da, db + lo_off, idesc, accumulate);
              umma_tf32_elect(d_tmem, da + lo_off, db, idesc, 1u);
              umma_tf32_elect(d_tmem, da, db, idesc, 1u);
              accumulate = 1;
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {  // 4 x (K = 8 tf32) per 32-channel chunk
              umma_tf32_elect(d_tmem, ((uint64_t)a_hi << 32) | (at + ks * 2), ((uint64_t)b_hi << 32) | (bt + ks * b_ks_step), idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit_elect(&empty_bar[stage]);  // smem slot is free once these MMAs have read it
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit_elect(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 6) {
    // ================= epilogue: TMEM -> registers -> (ReLU) -> NHWC rows =================
    const int sub = warp & 3;  // TMEM sub-partition this warp may read: lanes [32*sub, 32*sub+32)
    const int row = sub * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int tile = unit / p.nsplit, n0 = p.n0 + (unit - tile * p.nsplit) * p.N;
      const int img = p.ipt > 1 ? tile * p.ipt : tile / p.tiles_per_img, h0 = p.ipt > 1 ? 0 : (tile % p.tiles_per_img) * p.BH;
      const int rows_valid = p.ipt > 1 ? min(p.ipt, p.n_img - img) * p.Hg * p.Wg : min(p.BH, p.Hg - h0) * p.Wg;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      float* orow = p.out + ((size_t)(img * p.Hg + h0) * p.Wg + row) * p.ldc + n0;
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * p.N + c0), v);
        tmem_ld_wait();
        if (row < rows_valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (p.relu) {
              o.x = fmaxf(p.thr, o.x); o.y = fmaxf(p.thr, o.y); o.z = fmaxf(p.thr, o.z); o.w = fmaxf(p.thr, o.w);
            }
            *reinterpret_cast<float4*>(orow + c0 + j) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}


// =====================================================================================================================
// Halo implicit GEMM (fprop and dgrad, stride 1): every source pixel is fetched into shared memory ONCE per tile.
//   The tile is BH rows of the output grid.  ONE 4-D TMA box per 32-channel chunk brings the (BH+KH-1) x P source rows the
//   tile needs (P = Wg + KW - 1 pixels per row; pixels outside the source tensor -- dgrad's implicit zero padding, SAME
//   padding -- are zero-filled by TMA).  With the GEMM row index laid out on the SAME pitch, m = r*P + c, the source pixel
//   of filter tap (dh,dw) is row m + dh*P + dw of the shared tile: every tap's A operand is the one tile viewed through a
//   descriptor whose start address is shifted by (dh*P + dw) 128-byte rows (the 128-byte swizzle is a function of the
//   absolute shared-memory address, so TMA's write pattern and the MMA's read pattern agree for any row shift).
//   Rows with c >= Wg are computed and dropped by the epilogue (KW-1 of every P).  The whole filter is resident in shared
//   memory for the life of the CTA.  Per tile: kchunks TMA issues and taps*kchunks*4 MMAs -- no per-tap traffic.
// =====================================================================================================================
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all_groups() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct HaloParams {
  int n_tiles, tiles_per_img;
  int Hg, Wg, BH, P;
  int KH, KW, kchunks;
  int N;               // accumulator columns (fprop: Cout, dgrad: Cin), multiple of 32
  int b_mn_major;      // fprop 1 (B = N/32 boxes [32 n x 32 k]) ; dgrad 0 (B = one box [32 k x N rows])
  int b_rows_per_tap;  // rows between taps in the 2-D filter view
  int flip;            // dgrad: tap (kh,kw) reads source offset (KH-1-kh, KW-1-kw)
  int ox, oy;          // box origin relative to the tile's first output pixel (fprop: -PL,-PT ; dgrad: -(KW-1-PL), -(KH-1-PT))
  int a_tile_bytes;    // one 32-channel chunk of the source tile (multiple of 1024)
  int box_bytes;       // bytes TMA delivers per chunk
  int stages;
  int relu;
  float thr;
  float* out;          // [n_img, Hg, Wg, N]
  int tma_store;       // epilogue through swizzled shared staging + TMA stores (one box per grid row and 32-channel chunk)
  int early_filter;    // load the filter before the programmatic-dependency wait (SB_TC_EARLY_FILTER=0 turns it off)
  long long* dbg;      // developer timeline (SB_TC_DEBUG=1): 64 clock64 slots per CTA, null in production
  int bo_mode;         // debug: 1 = also encode the start row's swizzle phase in the descriptor's base-offset field
  // Fused Pool2D(2x2, stride 1, VALID) behind the (ReLU) epilogue: consecutive tiles overlap by one conv row (row_stride = BH - 1),
  // eight epilogue warps drain TMEM into the staging tile and then pool it: `pooled` [n_img, Hg-1, Wg-1, N] and one winner byte
  // per pooled output (kernels_fast.cu win4; `pmap` may be null) are written INSTEAD of the conv activations.
  int pool, row_stride;
  float* pooled;
  unsigned* pmap;
  // KW-folding (instantiations with a compile-time KW): the KW taps of one kernel row share ONE MMA -- the same A view (row shift of
  // the kernel row only) against B = [W(kh,0) | W(kh,1) | ...] with N = KW * N accumulator columns; column block j then holds the
  // tap's contribution displaced by j pixels (rows of the pitch-linear index), and the epilogue re-aligns:
  //   out[m] = sum_j D[m + d_j][block j],  d_j = j (fprop) or KW-1-j (dgrad, flipped taps).
  // At N = 64 an MMA is bound by its shared-memory operand fetch (4 KB of A + 32 B x N of B per K = 8 at 128 B/cycle); folding
  // fetches A once per kernel row instead of once per tap: 36 MMAs x 48 cycles -> 12 x 96 (fprop, N = 192, now math-bound), 72 x 40
  // -> 24 x 56 (dgrad, N = 96).
  int fold;
  // Output-channel slices (fprop, filters too large to stay resident whole): gridDim.y slices of N accumulator columns each; the CTA
  // keeps ITS slice of the filter resident and walks every tile.  ldc = channels of the output tensor (= N without slicing).
  int ldc;
};
// winner byte of a 2x2 window (the same definition as kernels_fast.cu win4): index of the FIRST maximum in row-major order | (max > thr) << 2
__device__ __forceinline__ unsigned tc_win4(float a, float b, float c, float d, int relu, float thr, float* mx) {
  float best = a;
  unsigned idx = 0;
  if (b > best) { best = b; idx = 1; }
  if (c > best) { best = c; idx = 2; }
  if (d > best) { best = d; idx = 3; }
  *mx = best;
  return idx | ((!relu || best > thr) ? 4u : 0u);
}

// TKH/TKW/TKC: compile-time KH, KW and Cin/32 (0 = from the parameters): runtime trip counts in the MMA issue loop cost far more
// than the MMAs themselves (profiles/r1_notes.md item 11)
// pooled epilogue: warps 2..9 drain TMEM, warps 2..25 (768 threads) pool -- with only the 8 drain warps pooling, the SM cannot hide
// the latency of the ~90-instruction item loop and a tile's pooling takes 3x its MMAs (measured: fprop 15 -> 23 us)
constexpr int kHaloPoolWarps = 24;
constexpr int kHaloPoolThreads = 64 + 32 * kHaloPoolWarps;
// FOLD: compile-time switch of the KW-folded variant (8 epilogue warps, 320 threads); the plain instantiation carries none of its code
template <int TKH, int TKW, int TKC, bool POOL = false, bool FOLD = false>
__global__ void __launch_bounds__(POOL ? kHaloPoolThreads : (FOLD ? kIgemmThreadsX3 : kIgemmThreads), 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int taps = p.KH * p.KW;
  const int b_tile_bytes = p.N * 128;                         // one (tap, chunk) slice of the filter
  const int b_bytes = taps * p.kchunks * b_tile_bytes;        // resident filter
  const int stage_bytes = p.kchunks * p.a_tile_bytes;
  uint8_t* sm_b = smem;
  uint8_t* sm_a = smem + b_bytes;
  uint8_t* sm_stg = sm_a + (size_t)p.stages * stage_bytes;        // 2 x (N/32) x [128 rows x 128 B], TMA-store staging
  const int stg_bytes = (p.tma_store || p.pool) ? (p.N / 32) * 16384 : 0;
  uint64_t* full_bar = (uint64_t*)(sm_stg + 2 * stg_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* b_bar = tmem_empty + 2;
  uint32_t* tmem_slot = (uint32_t*)(b_bar + 1);
  // KW-folding: rows a warp needs from the next TMEM sub-partition -- [2 tile parities][N/32 chunks][4 sub-partitions][KW-1 blocks][KW-1 rows][32]
  float* xch = reinterpret_cast<float*>(((uintptr_t)(tmem_slot + 4) + 15) & ~(uintptr_t)15);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // shfl: provably warp-uniform
  constexpr bool FOLD_OK = FOLD && !POOL && TKW >= 2 && TKW <= 4;  // folding needs a compile-time KW (register arrays in the epilogue)
  constexpr bool fold = FOLD_OK;
  constexpr int kEpiWarps = FOLD_OK ? 8 : 4, kEpiThreads = 32 * kEpiWarps;
  const int NB = fold ? TKW * p.N : p.N;  // accumulator columns per buffer
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * NB) tmem_cols <<= 1;
  long long* dbg = p.dbg && blockIdx.y == 0 ? p.dbg + (size_t)blockIdx.x * 64 : nullptr;
#define HALO_T(slot) do { if (dbg) dbg[slot] = clock64(); } while (0)
  const int n0 = (int)blockIdx.y * p.N;  // first output channel of this CTA's slice

  if (threadIdx.x == 0) {
    HALO_T(0);
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.tma_store) tma_prefetch_desc(&map_c);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], (POOL || FOLD_OK) ? 8 : 4);
    }
    mbar_init(b_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // The filter is only ever written by the optimizer / averaging / host uploads, never by the kernel right before this one (a
  // batch-gather or an activation kernel always sits in between, and that kernel itself waited for its predecessors), so its
  // load may start before the programmatic-dependency wait and overlap the previous kernel's tail.
  auto load_filter = [&]() {
    mbar_arrive_expect_tx(b_bar, (uint32_t)b_bytes);
    for (int tap = 0; tap < taps; ++tap)
      for (int kc = 0; kc < p.kchunks; ++kc) {
        // folded: [chunk][tap] so that the KW taps of a kernel row are adjacent for every channel chunk
        uint8_t* sb = sm_b + (size_t)(fold ? kc * taps + tap : tap * p.kchunks + kc) * b_tile_bytes;
        if (p.b_mn_major) {
          for (int j = 0; j < p.N / 32; ++j) tma_load_2d(sb + j * 4096, &map_b, b_bar, n0 + j * 32, tap * p.b_rows_per_tap + kc * 32);
        } else {
          tma_load_2d(sb, &map_b, b_bar, kc * 32, tap * p.b_rows_per_tap);
        }
      }
  };
  if (threadIdx.x == 0 && p.early_filter) load_filter();
  pdl_wait();  // the on-chip setup above overlapped the previous kernel; activations in global memory from here on
  pdl_trigger();  // dependents may start their own on-chip setup now (they still wait for this grid before touching its output)
  if (threadIdx.x == 0) HALO_T(1);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      if (!p.early_filter) load_filter();  // the filter, once
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int img = tile / p.tiles_per_img, h0 = (tile - img * p.tiles_per_img) * p.row_stride;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = sm_a + (size_t)stage * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(p.kchunks * p.box_bytes));
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_load_4d(sa + (size_t)kc * p.a_tile_bytes, &map_a, &full_bar[stage], kc * 32, p.ox, h0 + p.oy, img);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: warp-uniform loop, one elected lane issues =================
    {
      const uint32_t idesc = idesc_tf32(128, p.N, 0, p.b_mn_major);
      // descriptors: the high word is constant, the low word = (LBO >> 4) << 16 | (start address >> 4)
      const uint64_t a_d = smem_desc_sw128(0, 16, 1024), b_d = p.b_mn_major ? smem_desc(0, 4096, 512, 1) : smem_desc_sw128(0, 16, 1024);
      const uint32_t a_hi = (uint32_t)(a_d >> 32), a_lo0 = (uint32_t)a_d, b_hi = (uint32_t)(b_d >> 32), b_lo0 = (uint32_t)b_d;
      const uint32_t b_ks_step = p.b_mn_major ? (1024u >> 4) : (32u >> 4);
      const uint32_t sb0 = b_lo0 + (smem_u32(sm_b) >> 4);
      const uint32_t a_chunk = (uint32_t)(p.a_tile_bytes >> 4), b_chunk = (uint32_t)(b_tile_bytes >> 4);
      const int KH = TKH ? TKH : p.KH, KW = TKW ? TKW : p.KW, kchunks = TKC ? TKC : p.kchunks, P8 = p.P * 8;
      mbar_wait(b_bar, 0);
      tc_fence_after();
      if (lane == 0) HALO_T(2);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int ti = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++ti) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0 && ti < 6) HALO_T(8 + ti * 2);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NB);
        const uint32_t sa0 = a_lo0 + (smem_u32(sm_a + (size_t)stage * stage_bytes) >> 4);
        uint32_t accumulate = 0;
        uint32_t bt = sb0;
        if (fold) {
          // one MMA per (kernel row, chunk, k-step): A = the tile shifted by the kernel row only, B = the row's KW taps side by side
          const uint32_t idesc_f = idesc_tf32(128, NB, 0, p.b_mn_major);
          uint32_t row_at = sa0 + (p.flip ? (uint32_t)((KH - 1) * P8) : 0u);
          const int32_t step_h = p.flip ? -P8 : P8;
          const uint32_t b_row = (uint32_t)KW * b_chunk, b_kc = (uint32_t)(KH * KW) * b_chunk;  // [chunk][tap] layout
#pragma unroll
          for (int kh = 0; kh < KH; ++kh, row_at += step_h) {
            uint32_t at = row_at, btk = sb0 + (uint32_t)kh * b_row;
#pragma unroll
            for (int kc = 0; kc < kchunks; ++kc, at += a_chunk, btk += b_kc) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                umma_tf32_elect(d_tmem, ((uint64_t)a_hi << 32) | (at + ks * 2), ((uint64_t)b_hi << 32) | (btk + ks * b_ks_step), idesc_f,
                                accumulate);
                accumulate = 1;
              }
            }
          }
          umma_commit_elect(&empty_bar[stage]);
          umma_commit_elect(&tmem_full[acc]);
          if (lane == 0 && ti < 6) HALO_T(9 + ti * 2);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          continue;
        }
        // tap (kh,kw) reads source rows shifted by (dh*P + dw) pixels = 8 descriptor units (128 B >> 4) each
        uint32_t row_at = sa0 + (p.flip ? (uint32_t)((KH - 1) * P8 + (KW - 1) * 8) : 0u);
        const int32_t step_w = p.flip ? -8 : 8, step_h = p.flip ? -P8 : P8;
#pragma unroll
        for (int kh = 0; kh < KH; ++kh, row_at += step_h) {
          uint32_t at_tap = row_at;
#pragma unroll
          for (int kw = 0; kw < KW; ++kw, at_tap += step_w) {
            uint32_t at = at_tap;
#pragma unroll
            for (int kc = 0; kc < kchunks; ++kc, at += a_chunk, bt += b_chunk) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {  // 4 x (K = 8 tf32) per 32-channel chunk
                umma_tf32_elect(d_tmem, ((uint64_t)a_hi << 32) | (at + ks * 2), ((uint64_t)b_hi << 32) | (bt + ks * b_ks_step), idesc,
                                accumulate);
                accumulate = 1;
              }
            }
          }
        }
        umma_commit_elect(&empty_bar[stage]);  // the source tile is free once these MMAs have read it
        umma_commit_elect(&tmem_full[acc]);
        if (lane == 0 && ti < 6) HALO_T(9 + ti * 2);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (POOL) {
    // ================= pooled epilogue: TMEM -> (ReLU) -> staging tile -> 2x2/stride-1 max + winner byte =================
    // warps 2..5 drain the even 32-channel chunks, warps 6..9 the odd ones (a warp may only read the TMEM lanes of sub-partition
    // warp % 4); after one barrier all kHaloPoolWarps warps pool channel quads of the staged tile: window (r, c) = staged rows
    // m, m+1, m+P, m+P+1 (m = r*P + c), members in row-major order.  Items advance by a fixed stride, so (r, c) are updated
    // incrementally (no divisions in the loop).
    constexpr int NT = 32 * kHaloPoolWarps;
    const bool drains = warp < 10;
    const int sub = warp & 3, grp = (warp - 2) >> 2;
    const int m = sub * 32 + lane;
    const int et = threadIdx.x - 64;  // 0 .. NT-1
    const int PH = p.Hg - 1, PW = p.Wg - 1, C4 = p.N >> 2, nchunks = p.N >> 5;
    // item i = px * C4 + cq ; this thread's items: et, et + NT, ...  ->  px advances by dpx (+1 when cq wraps)
    const int cq0 = et % C4, px0 = et / C4, dcq = NT % C4, dpx = NT / C4;
    int acc = 0;
    uint32_t acc_phase = 0;
    int ti = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++ti) {
      const int img = tile / p.tiles_per_img, h0 = (tile - img * p.tiles_per_img) * p.row_stride;
      uint8_t* sbuf = sm_stg + (size_t)(ti & 1) * stg_bytes;
      if (drains) {
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        if (threadIdx.x == 64 && ti < 6) HALO_T(24 + ti * 2);
        for (int ch = grp; ch < nchunks; ch += 2) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * p.N + ch * 32), v);
          tmem_ld_wait();
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(p.thr, v[j]);
          }
          // row m of the chunk tile: 128 B, 16-byte piece j at position j ^ (m & 7) (conflict-free row writes and quad reads)
          uint8_t* srow = sbuf + (size_t)ch * 16384 + m * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(srow + ((j ^ (m & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        // the accumulator is drained: release it to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      }
      if (threadIdx.x == 64 && ti < 6) HALO_T(36 + ti);
      asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
      if (threadIdx.x == 64 && ti < 6) HALO_T(42 + ti);
      const int rows_pool = min(p.row_stride, PH - h0);
      const int npx = rows_pool * PW;
      float* pout = p.pooled + ((size_t)(img * PH + h0) * PW) * p.N;
      unsigned* mout = p.pmap ? p.pmap + ((size_t)(img * PH + h0) * PW) * C4 : nullptr;
      int cq = cq0, px = px0;
      int r = px / PW, c = px - r * PW;  // once per tile
      while (px < npx) {
        const int m0 = r * p.P + c;
        const uint8_t* base = sbuf + (size_t)(cq >> 3) * 16384;
        const int j = cq & 7;
        const float4 a = *reinterpret_cast<const float4*>(base + m0 * 128 + ((j ^ (m0 & 7)) << 4));
        const float4 b = *reinterpret_cast<const float4*>(base + (m0 + 1) * 128 + ((j ^ ((m0 + 1) & 7)) << 4));
        const float4 cc = *reinterpret_cast<const float4*>(base + (m0 + p.P) * 128 + ((j ^ ((m0 + p.P) & 7)) << 4));
        const float4 d = *reinterpret_cast<const float4*>(base + (m0 + p.P + 1) * 128 + ((j ^ ((m0 + p.P + 1) & 7)) << 4));
        float4 mx;
        const unsigned b0 = tc_win4(a.x, b.x, cc.x, d.x, p.relu, p.thr, &mx.x);
        const unsigned b1 = tc_win4(a.y, b.y, cc.y, d.y, p.relu, p.thr, &mx.y);
        const unsigned b2 = tc_win4(a.z, b.z, cc.z, d.z, p.relu, p.thr, &mx.z);
        const unsigned b3 = tc_win4(a.w, b.w, cc.w, d.w, p.relu, p.thr, &mx.w);
        *reinterpret_cast<float4*>(pout + (size_t)px * p.N + cq * 4) = mx;
        if (mout) mout[(size_t)px * C4 + cq] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        cq += dcq; px += dpx; c += dpx;
        if (cq >= C4) { cq -= C4; ++px; ++c; }
        while (c >= PW) { c -= PW; ++r; }
      }
      // (no second barrier: the staging buffers alternate, and a thread reaches the NEXT tile's barrier only after this loop)
      if (threadIdx.x == 64 && ti < 6) HALO_T(25 + ti * 2);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ================= epilogue: TMEM -> registers -> (ReLU) -> NHWC rows =================
    // (folded variant: eight warps -- warps 2..5 own the even 32-channel chunks, warps 6..9 the odd ones)
    const int grp = (warp - 2) >> 2;
    const int sub = warp & 3;  // TMEM sub-partition this warp may read: lanes [32*sub, 32*sub+32)
    const int m = sub * 32 + lane;
    const int r = m / p.P, c = m - r * p.P;
    int acc = 0;
    uint32_t acc_phase = 0;
    int ti = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++ti) {
      const int img = tile / p.tiles_per_img, h0 = (tile - img * p.tiles_per_img) * p.row_stride;
      const bool valid = r < p.BH && c < p.Wg && h0 + r < p.Hg;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (threadIdx.x == 64 && ti < 6) HALO_T(24 + ti * 2);
      float* orow = p.out + ((size_t)(img * p.Hg + h0 + r) * p.Wg + c) * p.ldc + n0;
      uint8_t* sbuf = sm_stg + (size_t)(ti & 1) * stg_bytes;
      if (p.tma_store) {
        // the staging buffer is free once the stores issued two tiles ago have READ it
        if (threadIdx.x == 64) tma_store_wait_read1();
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
      const int nch = p.N >> 5;
      const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * NB);
      if (FOLD_OK && fold) {
        // pass 1: the first KW-1 rows of every warp are the rows its predecessor's last lanes will need (displaced blocks only)
        for (int ch = grp; ch < nch; ch += 2)
#pragma unroll
          for (int d = 1; d < (FOLD_OK ? TKW : 1); ++d) {
            const int jb = p.flip ? TKW - 1 - d : d;
            float t[32];
            tmem_ld32(t_row + (uint32_t)(jb * p.N + ch * 32), t);
            tmem_ld_wait();
            if (lane < d) {
              float* dst = xch + ((((size_t)((ti & 1) * nch + ch) * 4 + sub) * (TKW - 1) + (d - 1)) * (TKW - 1) + lane) * 32;
#pragma unroll
              for (int q = 0; q < 32; q += 4) *reinterpret_cast<float4*>(dst + q) = make_float4(t[q], t[q + 1], t[q + 2], t[q + 3]);
            }
          }
        asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory");
      }
      for (int c0 = (FOLD_OK ? grp * 32 : 0); c0 < p.N; c0 += (FOLD_OK ? 64 : 32)) {
        float v[32];
        if (FOLD_OK && fold) {
          // out[m] = D[m][block of displacement 0] + D[m+1][displacement 1] + ... : the other rows live in the next lanes (shuffle)
          // or, for the last lanes of a warp, in the next sub-partition (pass 1's exchange rows; rows >= 128 only feed dropped outputs)
          tmem_ld32(t_row + (uint32_t)((p.flip ? TKW - 1 : 0) * p.N + c0), v);
          tmem_ld_wait();
#pragma unroll
          for (int d = 1; d < (FOLD_OK ? TKW : 1); ++d) {
            const int jb = p.flip ? TKW - 1 - d : d;
            float t[32];
            tmem_ld32(t_row + (uint32_t)(jb * p.N + c0), t);
            tmem_ld_wait();
            const bool over = lane + d >= 32;
            // lane i takes lane i+d's row; for the last d lanes that row lives in the next sub-partition: pass 1's exchange rows
            // (their OWN registers stay untouched -- they are the shuffle sources of the lanes below them)
            float xr[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) xr[q] = 0.f;
            if (over && sub < 3) {
              const float* src = xch + ((((size_t)((ti & 1) * nch + (c0 >> 5)) * 4 + (sub + 1)) * (TKW - 1) + (d - 1)) * (TKW - 1) +
                                        ((lane + d) & 31)) * 32;
#pragma unroll
              for (int q = 0; q < 32; q += 4) {
                const float4 x4 = *reinterpret_cast<const float4*>(src + q);
                xr[q] = x4.x; xr[q + 1] = x4.y; xr[q + 2] = x4.z; xr[q + 3] = x4.w;
              }
            }
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              const float up = __shfl_down_sync(0xffffffffu, t[q], d);
              v[q] = __fadd_rn(v[q], over ? xr[q] : up);
            }
          }
        } else {
          tmem_ld32(t_row + (uint32_t)c0, v);
          tmem_ld_wait();
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(p.thr, v[j]);
        }
        if (p.tma_store) {
          // row m of the 32-channel chunk tile: 128 B, 16-byte piece j at position j ^ (m & 7) (the 128B swizzle of map_c)
          uint8_t* srow = sbuf + (size_t)(c0 >> 5) * 16384 + m * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(srow + ((j ^ (m & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
      // the accumulator is drained: release it to the MMA warp before the (slower) store issue
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (p.tma_store) {
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        if (threadIdx.x == 64) {
          const int rows = min(p.BH, p.Hg - h0);
          for (int rr = 0; rr < rows; ++rr)
            for (int cc = 0; cc < p.N / 32; ++cc)
              tma_store_4d(&map_c, sbuf + (size_t)cc * 16384 + (size_t)rr * p.P * 128, n0 + cc * 32, 0, h0 + rr, img);
          tma_store_commit_group();
        }
      }
      if (threadIdx.x == 64 && ti < 6) HALO_T(25 + ti * 2);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_store && threadIdx.x == 64) tma_store_wait_all_groups();  // global writes complete before the kernel ends
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) HALO_T(3);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
#undef HALO_T
}

// =====================================================================================================================
// Conv2D wgrad:  dW[tap][ci][co] = sum_pixels X[n, oh+kh, ow+kw, ci] * dY[n, oh, ow, co]
//   D_tap[M = co, N = ci] += A[M, K] . B_tap[K, N],  K = pixels.  A = dY rows (MN-major, M = Cout contiguous),
//   B_tap = X rows shifted by the tap (MN-major, N = Cin contiguous).  K-chunk = BH grid rows x WK pixels (WK = OW rounded up
//   to 8; the extra columns of dY are TMA zero fill, so they add nothing).  One CTA accumulates its chunks for ALL taps in
//   TMEM (taps*Cin columns), then writes one partial filter gradient; partials are reduced in CTA order.
// =====================================================================================================================
struct WgradParams {
  int n_chunks, chunks_per_img;
  int BH, WK;
  int KH, KW, Cin, Cout;
  // blockIdx.y picks a slice of the filter: taps [t0, t0 + nt) (nt * Cin <= 512 TMEM columns) x output channels [co0, co0 + Mh);
  // blockIdx.x is the K split (the chunks ch = blockIdx.x, + gridDim.x, ...).  One launch covers every slice, so the number of
  // partial filter gradients is the K split (148 / slices), not 148 per slice.
  int n_slices, nt_max;
  unsigned char s_t0[16], s_nt[16];
  unsigned short s_co0[16];
  int Mh;          // 64 or 128 (the UMMA M)
  int pt, pl;      // padding before: tap (kh,kw) pairs dY(oh,ow) with X(oh + kh - pt, ow + kw - pl); OOB pixels are zero-filled
  int stages;
  float* partial;  // [gridDim.x][KH*KW*Cin*Cout]
  long long* dbg;  // developer timeline (SB_TC_DEBUG=1): 64 clock64 slots per CTA, null in production
  int x3;          // 3xTF32: converter warps write the low-order tiles behind the stage's [dY | X taps]
};
constexpr int kWgradThreads = 192;
constexpr int kWgradMaxSlices = 16;

__global__ void __launch_bounds__(kIgemmThreadsX3, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  long long* dbg = p.dbg && blockIdx.y == 0 ? p.dbg + (size_t)blockIdx.x * 64 : nullptr;
#define WG_T(slot) do { if (dbg) dbg[slot] = clock64(); } while (0)
  if (threadIdx.x == 0) WG_T(0);
  const int taps = p.s_nt[blockIdx.y], t0 = p.s_t0[blockIdx.y], co0 = p.s_co0[blockIdx.y];  // this CTA's slice of the filter
  const int krows = p.BH * p.WK;            // K rows per chunk (multiple of 8)
  const int box_bytes = krows * 128;        // one 32-channel box
  const int a_bytes = (p.Mh / 32) * box_bytes;
  const int b_bytes = taps * (p.Cin / 32) * box_bytes;
  const int hi_bytes = ((a_bytes + p.nt_max * (p.Cin / 32) * box_bytes + 1023) / 1024) * 1024;  // stage layout of the widest slice
  const int stage_bytes = p.x3 ? 2 * hi_bytes : hi_bytes;
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* conv_bar = empty_bar + p.stages;  // 3xTF32: the stage's low-order tiles are written
  uint64_t* done_bar = conv_bar + p.stages;
  uint32_t* tmem_slot = (uint32_t*)(done_bar + 1);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // shfl: provably warp-uniform
  const int ncols = taps * p.Cin;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < ncols) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&conv_bar[s], 4);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) WG_T(1);
  pdl_wait();  // the on-chip setup above overlapped the previous kernel; global memory from here on
  pdl_trigger();  // dependents may start their own on-chip setup now (they still wait for this grid before touching its output)
  if (threadIdx.x == 0) WG_T(2);
  const bool has_work = (int)blockIdx.x < p.n_chunks;

  if (warp >= 6) {
    // 3xTF32 converters (launched only in that mode): lo = x - tf32(x) for the stage's dY and X boxes
    const int t = threadIdx.x - 192;
    int stage = 0;
    uint32_t phase = 0;
    for (int ch = blockIdx.x; ch < p.n_chunks; ch += gridDim.x) {
      mbar_wait(&full_bar[stage], phase);
      uint8_t* st = smem + (size_t)stage * stage_bytes;
      split_lo_tile(st, st + hi_bytes, a_bytes + b_bytes, t);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&conv_bar[stage]);
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ch = blockIdx.x; ch < p.n_chunks; ch += gridDim.x) {
        const int img = ch / p.chunks_per_img, h0 = (ch % p.chunks_per_img) * p.BH;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(a_bytes + b_bytes));
        for (int j = 0; j < p.Mh / 32; ++j) tma_load_4d(sa + j * box_bytes, &map_dy, &full_bar[stage], co0 + j * 32, 0, h0, img);
        for (int t = 0; t < taps; ++t)
          for (int j = 0; j < p.Cin / 32; ++j)
            tma_load_4d(sb + (t * (p.Cin / 32) + j) * box_bytes, &map_x, &full_bar[stage], j * 32, (t0 + t) % p.KW - p.pl,
                        h0 + (t0 + t) / p.KW - p.pt, img);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (has_work) {  // warp-uniform loop, one elected lane issues
      const uint32_t idesc = idesc_tf32(p.Mh, p.Cin, 1, 1);  // M = 64 or 128 output channels, N = Cin, both MN-major
      // MN-major SW128 (32-byte atoms): 8 K-rows per instruction (one swizzle atom deep), 32-element groups along M/N LBO apart
      const uint64_t tmpl = smem_desc(0, (uint32_t)box_bytes, 512, 1);
      const uint32_t d_hi = (uint32_t)(tmpl >> 32), d_lo = (uint32_t)tmpl;
      const uint32_t b_tap = (uint32_t)(((p.Cin / 32) * box_bytes) >> 4);
      const int nks = krows / 8;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accumulate = 0;
      int ti = 0;
      for (int ch = blockIdx.x; ch < p.n_chunks; ch += gridDim.x, ++ti) {
        mbar_wait(p.x3 ? &conv_bar[stage] : &full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0 && ti < 12) WG_T(8 + 2 * ti);
        const uint32_t sa = d_lo + (smem_u32(smem + (size_t)stage * stage_bytes) >> 4);
        uint32_t sb = sa + (uint32_t)(a_bytes >> 4);
        const uint32_t lo_off = (uint32_t)(hi_bytes >> 4);
        for (int t = 0; t < taps; ++t, sb += b_tap) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(t * p.Cin);
          uint32_t acc_t = accumulate;
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t da = ((uint64_t)d_hi << 32) | (sa + ks * 64), db = ((uint64_t)d_hi << 32) | (sb + ks * 64);
            if (p.x3) {  // small terms first: A_hi.B_lo, A_lo.B_hi, then A_hi.B_hi
              umma_tf32_elect(d_tmem, da, db + lo_off, idesc, acc_t);
              umma_tf32_elect(d_tmem, da + lo_off, db, idesc, 1u);
              umma_tf32_elect(d_tmem, da, db, idesc, 1u);
            } else {
              umma_tf32_elect(d_tmem, da, db, idesc, acc_t);
            }
            acc_t = 1;
          }
        }
        accumulate = 1;
        umma_commit_elect(&empty_bar[stage]);
        if (lane == 0 && ti < 12) WG_T(9 + 2 * ti);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit_elect(done_bar);
    }
  } else if (warp < 6) {
    // epilogue: partial[cta][(t*Cin + ci)*Cout + co] = D_t[co][ci]
    const int sub = warp & 3;
    float* part = p.partial + (size_t)blockIdx.x * p.KH * p.KW * p.Cin * p.Cout + (size_t)t0 * p.Cin * p.Cout + co0;
    int co = -1;
    if (p.Mh == 128) co = sub * 32 + lane;                    // M = 128: TMEM lane == row
    else if (lane < 16) co = sub * 16 + lane;                 // M = 64: row m sits in lane (m % 16) + 32 * (m / 16)
    if (has_work) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
    if (threadIdx.x == 64) WG_T(3);
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      float v[32];
      if (has_work) {
        tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (co >= 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) part[(size_t)(c0 + j) * p.Cout + co] = v[j];
      }
    }
    if (threadIdx.x == 64) WG_T(4);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) WG_T(5);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
#undef WG_T
}


// =====================================================================================================================
// Conv2D wgrad, halo form (stride 1).  K = pixels of a chunk of BH output rows laid out on a pitch of P pixels (k = r*P + c);
// per chunk ONE box brings the dY rows [32 co, P, BH] (columns >= OW zero-filled) and ONE box the X halo rows
// [32 ci, P, BH+KH-1] per 32-channel slice.  For a kernel row kh and channel slice j the A operand is X^T viewed MN-major with
// its 32-element M groups ONE pixel row (128 B) apart (LBO = 128): group g is the tap kw = g, so a single M = 128 MMA computes
// the KW taps of the row at once (rows 32*KW.. of D are unused):
//     D_(kh,j)[m = (kw, ci), n = co] += sum_k X[k + kh*P + kw][ci] * dY[k][co]
// -- KH * Cin/32 MMAs per 8 pixels instead of KH*KW*Cin/32.  Accumulators for all (kh, j) stay in TMEM for the CTA's chunks;
// one partial filter gradient per CTA, reduced in CTA order afterwards (deterministic).
// =====================================================================================================================
struct WgradHaloParams {
  int n_chunks, chunks_per_img;
  int BH, P, KH, KW, Cin, Cout, OH;
  int Ns;               // output channels per CTA (the MMA N): blockIdx.y picks the slice [blockIdx.y * Ns, + Ns); KH * Cin/32 * Ns <= 512 TMEM columns
  int ox, oy;           // X box origin relative to the chunk's first output pixel: (-PL, -PT)
  int dy_tile_bytes;    // one 32-co slice of the dY chunk (multiple of 1024)
  int x_tile_bytes;     // one 32-ci slice of the X halo chunk, incl. the KW-1 zeroed tail rows (multiple of 1024)
  int x_box_bytes, dy_box_bytes;
  int stages;
  float* partial;       // [gridDim.x][KH*KW*Cin*Cout]
  long long* dbg;       // developer timeline (SB_TC_DEBUG=1)
  int dbg_mode;         // developer experiment (SB_WG_EXPERIMENT): 1 = no TMA after the first fill of every stage (wrong results)
};

// TKH / TCIS: compile-time KH and Cin/32 (0 = read them from the parameters) so that the MMA issue loop is fully unrolled
template <int TKH, int TCIS>
__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WgradHaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  long long* dbg = p.dbg && blockIdx.y == 0 ? p.dbg + (size_t)blockIdx.x * 64 : nullptr;
#define WG_T(slot) do { if (dbg) dbg[slot] = clock64(); } while (0)
  if (threadIdx.x == 0) WG_T(0);
  const int KH = TKH ? TKH : p.KH;
  const int Ns = p.Ns, co0 = (int)blockIdx.y * Ns;
  const int co_slices = Ns / 32, ci_slices = TCIS ? TCIS : p.Cin / 32;
  const int a_bytes = ci_slices * p.x_tile_bytes, b_bytes = co_slices * p.dy_tile_bytes;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* done_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = (uint32_t*)(done_bar + 1);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int n_acc = KH * ci_slices;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < n_acc * Ns) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  // rows of every X tile that TMA never writes (they meet zero dY rows) must not hold NaN bit patterns
  for (int s = 0; s < p.stages; ++s)
    for (int j = 0; j < ci_slices; ++j) {
      uint8_t* t = smem + (size_t)s * stage_bytes + (size_t)j * p.x_tile_bytes;
      for (int i = p.x_box_bytes + threadIdx.x * 16; i < p.x_tile_bytes; i += blockDim.x * 16) *reinterpret_cast<float4*>(t + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) WG_T(1);
  pdl_wait();  // the on-chip setup above overlapped the previous kernel; global memory from here on
  pdl_trigger();  // dependents may start their own on-chip setup now (they still wait for this grid before touching its output)
  if (threadIdx.x == 0) WG_T(2);
  const bool has_work = (int)blockIdx.x < p.n_chunks;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ch = blockIdx.x; ch < p.n_chunks; ch += gridDim.x) {
        const int img = ch / p.chunks_per_img, h0 = (ch - img * p.chunks_per_img) * p.BH;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        if (p.dbg_mode == 1 && ch >= (int)blockIdx.x + p.stages * (int)gridDim.x) {
          mbar_arrive(&full_bar[stage]);
        } else {
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(ci_slices * p.x_box_bytes + co_slices * p.dy_box_bytes));
          for (int j = 0; j < ci_slices; ++j) tma_load_4d(sa + (size_t)j * p.x_tile_bytes, &map_x, &full_bar[stage], j * 32, p.ox, h0 + p.oy, img);
          for (int j = 0; j < co_slices; ++j) tma_load_4d(sb + (size_t)j * p.dy_tile_bytes, &map_dy, &full_bar[stage], co0 + j * 32, 0, h0, img);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (has_work) {  // warp-uniform loop, one elected lane issues
      const uint32_t idesc = idesc_tf32(128, Ns, 1, 1);
      // MN-major SW128 (32-byte atoms): 8 K-rows (1024 B) per instruction; A: 32-element M groups 128 B apart (the KW taps);
      // B: 32-element N groups one dY slice apart
      const uint64_t a_d = smem_desc(0, 128, 512, 1), b_d = smem_desc(0, (uint32_t)p.dy_tile_bytes, 512, 1);
      const uint32_t a_hi = (uint32_t)(a_d >> 32), a_lo0 = (uint32_t)a_d, b_hi = (uint32_t)(b_d >> 32), b_lo0 = (uint32_t)b_d;
      const int nks = p.BH * p.P / 8;
      const uint32_t kh_step = (uint32_t)(p.P * 128) >> 4, x_tile = (uint32_t)p.x_tile_bytes >> 4;
      // the TMEM base came through shared memory into a vector register: the shuffle makes it provably uniform, so the
      // accumulator addresses stay in uniform registers (otherwise every MMA pays a VIADD + R2UR: 118 vs 48 cycles per MMA)
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accumulate = 0;
      int ti = 0;
      for (int ch = blockIdx.x; ch < p.n_chunks; ch += gridDim.x, ++ti) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0 && ti < 12) WG_T(8 + 2 * ti);
        const uint32_t sa = a_lo0 + (smem_u32(smem + (size_t)stage * stage_bytes) >> 4);
        const uint32_t sb = b_lo0 + ((smem_u32(smem + (size_t)stage * stage_bytes) + (uint32_t)a_bytes) >> 4);
        for (int ks = 0; ks < nks; ++ks) {
          uint32_t d_tmem = tmem_u;
          uint32_t at_kh = sa + ks * 64;
#pragma unroll
          for (int kh = 0; kh < KH; ++kh, at_kh += kh_step) {
            uint32_t at = at_kh;
#pragma unroll
            for (int j = 0; j < ci_slices; ++j, at += x_tile, d_tmem += Ns)
              umma_tf32_elect(d_tmem, ((uint64_t)a_hi << 32) | at, ((uint64_t)b_hi << 32) | (sb + ks * 64), idesc, accumulate);
          }
          accumulate = 1;
        }
        umma_commit_elect(&empty_bar[stage]);
        if (lane == 0 && ti < 12) WG_T(9 + 2 * ti);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit_elect(done_bar);
    }
  } else {
    // epilogue: partial[cta][((kh*KW + kw)*Cin + j*32 + ci)*Cout + co] = D_(kh,j)[kw*32 + ci][co]
    const int sub = warp & 3;  // = kw (TMEM lanes 32*sub .. 32*sub+31)
    float* part = p.partial + (size_t)blockIdx.x * p.KH * p.KW * p.Cin * p.Cout;
    if (has_work) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
    if (threadIdx.x == 64) WG_T(3);
    // This warp's 32 rows (one kw, 32 ci) of accumulator a are ONE contiguous 32 x Cout block of the partial.  Each lane holds a
    // row: transpose through a padded scratch (the pipeline stages are idle once done_bar fired; row pitch Cout + 4 floats keeps
    // both the 128-bit row writes and the 128-bit line reads conflict-free) and store whole 512-byte pieces.
    const int pitch = Ns + 4;
    const uint32_t scratch = smem_u32(smem) + (uint32_t)(sub * 32 * pitch * 4);
    const int f4_per_row = Ns / 4;                // Ns % 32 == 0
    const int rows_per_it = 32 / f4_per_row > 0 ? 32 / f4_per_row : 1;
    for (int a = 0; a < n_acc; ++a) {
      const int kh = a / ci_slices, j = a - kh * ci_slices;
      float* blk = part + ((size_t)((kh * p.KW + sub) * p.Cin + j * 32)) * p.Cout + co0;
      for (int c0 = 0; c0 < Ns; c0 += 32) {
        float v[32];
        if (has_work) {
          tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(a * Ns + c0), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = 0.f;
        }
        const uint32_t dst = scratch + (uint32_t)((lane * pitch + c0) * 4);
#pragma unroll
        for (int q = 0; q < 32; q += 4)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + q * 4), "f"(v[q]), "f"(v[q + 1]), "f"(v[q + 2]), "f"(v[q + 3]) : "memory");
      }
      __syncwarp();
      if (sub < p.KW) {
        if (f4_per_row <= 32) {
          const int rr = lane / f4_per_row, cc = lane - rr * f4_per_row;
          for (int r0 = 0; r0 < 32; r0 += 4 * rows_per_it) {
            float4 t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + u * rows_per_it + rr;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t[u].x), "=f"(t[u].y), "=f"(t[u].z), "=f"(t[u].w)
                           : "r"(scratch + (uint32_t)((r * pitch + cc * 4) * 4)));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + u * rows_per_it + rr;
              *reinterpret_cast<float4*>(blk + (size_t)r * p.Cout + cc * 4) = t[u];
            }
          }
        } else {
          for (int r = 0; r < 32; ++r)
            for (int c = lane; c < f4_per_row; c += 32) {
              float4 t;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                           : "r"(scratch + (uint32_t)((r * pitch + c * 4) * 4)));
              *reinterpret_cast<float4*>(blk + (size_t)r * p.Cout + c * 4) = t;
            }
        }
      }
      __syncwarp();
    }
    if (threadIdx.x == 64) WG_T(4);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) WG_T(5);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
#undef WG_T
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int parts) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = partial[i];
  for (int c = 1; c < parts; ++c) acc = __fadd_rn(acc, partial[(size_t)c * n + i]);
  out[i] = acc;
}

// =====================================================================================================================
// per-layer plans
// =====================================================================================================================
struct DenseTcPlan {
  TcGemm* fwd = nullptr;    // Z = X . W (+bias, activation in the epilogue)
  TcGemm* dgrad = nullptr;  // dX = G . W^T
  TcGemm* wgrad = nullptr;  // dW = X^T . G
};

struct ConvTcPlan {
  CUtensorMap map_x_fwd, map_w_fwd;    // fprop: A = X boxes, B = filter (MN-major)
  CUtensorMap map_dy_dg, map_w_dg;     // dgrad: A = dY boxes, B = filter (K-major)
  CUtensorMap map_dy_wg, map_x_wg;     // wgrad
  IgemmParams fwd{}, dg{};
  HaloParams hfwd{}, hdg{};
  CUtensorMap map_x_h, map_dy_h;       // halo kernels: source boxes [32 ch, P, BH+KH-1, 1]
  CUtensorMap map_y_out, map_dx_out;   // halo kernels: TMA-store boxes [32 ch, Wg, 1, 1]
  bool halo_fwd = false, halo_dg = false;
  int hfwd_smem = 0, hdg_smem = 0, hfwd_grid = 0, hdg_grid = 0, hfwd_slices = 1;
  WgradParams wg{};
  std::vector<WgradParams> wg_list;   // launches of the per-tap wgrad kernel (tap groups x channel halves)
  bool wg_ok = false;
  WgradHaloParams hwg{};
  CUtensorMap map_dy_hwg, map_x_hwg;
  bool halo_wg = false;
  int hwg_smem = 0, hwg_grid = 0, hwg_slices = 1;
  int fwd_smem = 0, dg_smem = 0, wg_smem = 0;
  int fwd_grid = 0, dg_grid = 0, wg_grid = 0;
  float* wg_partial = nullptr;
  bool relu_fused = false;
  bool has_dgrad = false;
};

static bool conv_fits(const ConvGeom& g) {
  // stride 1 (any padding: TMA zero-fills out-of-bounds pixels), channel counts that tile into 32-float (128 B) swizzle rows
  // and legal UMMA N.  wgrad additionally needs Cout % 64 == 0 (UMMA M) and runs as several launches when taps*Cin > 512
  // TMEM columns or Cout > 128; when it does not fit the fp32 kernel computes the filter gradient.
  if (g.SH != 1 || g.SW != 1) return false;
  if (g.Cin % 32 || g.Cout % 32 || g.Cin > 256 || g.Cout > 256) return false;
  if (g.OW > 128 || g.W > 128) return false;
  return true;
}


// Fill a HaloParams for an output grid Hg x Wg whose taps read a source tensor [Csrc, Ws, Hs, N] at origin (ox, oy).
static bool plan_halo(HaloParams& p, CUtensorMap* map, CUtensorMap* map_out, const float* src, int Csrc, int Ws, int Hs, int Nimg, int Hg, int Wg,
                      int KH, int KW, int Nfull, int b_mn_major, int b_rows_per_tap, int flip, int ox, int oy, float* out,
                      int dev_smem, int* smem_out, int pool = 0, int n_slices = 1) {
  if (n_slices < 1 || Nfull % (32 * n_slices) || (n_slices > 1 && (pool || !b_mn_major))) return false;  // slices: fprop only
  const int Nacc = Nfull / n_slices;
  p.Hg = Hg; p.Wg = Wg; p.KH = KH; p.KW = KW; p.kchunks = Csrc / 32; p.N = Nacc; p.ldc = Nfull;
  p.P = Wg + KW - 1;
  if (p.P > 128) return false;
  p.BH = std::max(1, std::min(128 / p.P, Hg));
  const int box_h = p.BH + KH - 1;
  if (p.P > 256 || box_h > 256) return false;
  p.pool = pool; p.pooled = nullptr; p.pmap = nullptr;
  if (pool) {  // tiles overlap by one conv row: tile t pools rows [t*(BH-1), t*(BH-1) + BH-1) out of conv rows [t*(BH-1), t*(BH-1) + BH)
    if (p.BH < 2 || Hg < 2 || Wg < 2) return false;
    p.row_stride = p.BH - 1;
    p.tiles_per_img = (Hg - 1 + p.row_stride - 1) / p.row_stride;
  } else {
    p.row_stride = p.BH;
    p.tiles_per_img = (Hg + p.BH - 1) / p.BH;
  }
  p.n_tiles = Nimg * p.tiles_per_img;
  p.b_mn_major = b_mn_major; p.b_rows_per_tap = b_rows_per_tap; p.flip = flip; p.ox = ox; p.oy = oy;
  const int rows = std::max(box_h * p.P, 128 + (KH - 1) * p.P + KW - 1);
  p.a_tile_bytes = ((rows + 7) / 8) * 1024;
  p.box_bytes = box_h * p.P * 128;
  p.relu = 0; p.thr = 0.f; p.out = out;
  p.dbg = nullptr;
  const char* dbg = getenv("SB_TC_DEBUG");
  if (dbg && dbg[0] == '1') {
    SB_CUDA(cudaMalloc((void**)&p.dbg, (size_t)kNumSMs * 64 * sizeof(long long)));
    SB_CUDA(cudaMemset(p.dbg, 0, (size_t)kNumSMs * 64 * sizeof(long long)));
  }
  { const char* ef = getenv("SB_TC_EARLY_FILTER"); p.early_filter = !(ef && ef[0] == '0'); }
  const char* bo = getenv("SB_TC_BASE_OFFSET");
  p.bo_mode = (bo && bo[0] == '1') ? 1 : 0;
  const int b_bytes = KH * KW * p.kchunks * Nacc * 128;
  const int stage_bytes = p.kchunks * p.a_tile_bytes;
  // KW-folding: the instantiations with compile-time KW = 3 (KH = 3, one or two channel chunks), accumulators 2 x KW*N <= 512 columns
  // Default: dgrad only (flip) -- there N = Cin is small, the unfolded MMAs are at their worst (40 cycles for N = 32) and the
  // re-aligning epilogue is light: LeNet conv2 dgrad 20.9 -> 18.7 us eager.  For fprop (N = 64 -> 192) the epilogue's 128 shuffles
  // per thread and tile outweigh the saved MMA cycles (15.0 -> 18.8 us); SB_TC_FOLD=2 folds both, =0 none.
  { const char* fo = getenv("SB_TC_FOLD");
    const int mode = fo ? atoi(fo) : 1;
    const bool geom = !pool && KH == 3 && KW == 3 && (p.kchunks == 1 || p.kchunks == 2) && KW * Nacc <= 256;
    p.fold = (geom && (mode >= 2 || (mode == 1 && flip))) ? 1 : 0; }
  const int xch_bytes = p.fold ? 2 * (Nacc / 32) * 4 * (KW - 1) * (KW - 1) * 128 + 32 : 0;
  const char* ts = getenv("SB_TC_TMA_STORE");
  p.tma_store = !pool && !(ts && ts[0] == '0') && p.Wg <= 256;
  int fixed = b_bytes + 16 * 16 + 1024 + 256 + xch_bytes + ((p.tma_store || pool) ? 2 * (Nacc / 32) * 16384 : 0);
  if (pool && (dev_smem - fixed) / stage_bytes < 2) return false;  // the pooled epilogue needs its staging tiles
  if (p.tma_store && (dev_smem - fixed) / stage_bytes < 2) {  // staging does not fit: per-thread stores
    p.tma_store = 0;
    fixed = b_bytes + 16 * 16 + 1024 + 256 + xch_bytes;
  }
  p.stages = std::min(6, (dev_smem - fixed) / stage_bytes);
  if (p.stages < 2) return false;
  {
    uint64_t od[4] = {(uint64_t)Nfull, (uint64_t)Wg, (uint64_t)Hg, (uint64_t)Nimg};
    uint32_t ob[4] = {32, (uint32_t)Wg, 1, 1};
    *map_out = make_map(out, 4, od, ob);
  }
  *smem_out = fixed + p.stages * stage_bytes;
  uint64_t d[4] = {(uint64_t)Csrc, (uint64_t)Ws, (uint64_t)Hs, (uint64_t)Nimg};
  uint32_t b[4] = {32, (uint32_t)p.P, (uint32_t)box_h, 1};
  *map = make_map(src, 4, d, b);
  return true;
}

// tiles of whole images when an image fills less than half of the MMA's 128 rows
static void igemm_small_grid(IgemmParams& p, int n_img) {
  p.n_img = n_img;
  p.ipt = 1;
  const int px = p.Hg * p.Wg;
  if (p.BH == p.Hg && px <= 64) {
    p.ipt = std::max(1, std::min(128 / px, n_img));
    p.n_tiles = (n_img + p.ipt - 1) / p.ipt;
  }
}
// N slices per tile (fprop): the cheaper of {1, 2, 4} under  max(per-CTA time, aggregate L2 -> shared-memory time), with a CTA pulling
// ~55 B/cycle, the L2 delivering ~3400 B/cycle to all SMs, and an M = 128 MMA of K = 8 taking max(N/2, 40) cycles (profiles/r1_notes.md item 8)
static int igemm_pick_nsplit(int n_tiles, int kblocks, int a_box_bytes, int n_full) {
  int best = 1;
  double best_cost = 1e30;
  for (int ns = 1; ns <= 4; ns *= 2) {
    if (n_full % (32 * ns) || n_full / ns < 64) break;
    const int n = n_full / ns;
    const double units = (double)n_tiles * ns, rounds = std::ceil(units / kNumSMs);
    const double bytes = (double)kblocks * (a_box_bytes + n * 128.0);
    const double mma = kblocks * 4.0 * std::max(n / 2.0, 40.0);
    const double cost = std::max(rounds * std::max(bytes / 55.0, mma), units * bytes / 3400.0);
    if (cost < best_cost * 0.95) { best_cost = cost; best = ns; }
  }
  return best;
}
static int igemm_smem(const IgemmParams& p) { return p.stages * (p.x3 ? 2 : 1) * (kATileBytes + p.N * 128) + 16 * (2 * p.stages + 4) + 16 + 1024; }

void tc_plan_layers(sb_engine* e) {
  if (e->train.precision == SB_PREC_FP32) return;
  // SB_PREC_3XTF32: every contraction runs with split operands (hi = tf32(x), lo = x - hi; A_hi.B_lo + A_lo.B_hi + A_hi.B_hi):
  // fp32-class accuracy on the tensor cores.  MatMul: gemm_tf32_kernel; Conv2D fprop / dgrad / wgrad: the per-tap kernels
  // (conv_igemm_kernel, conv_wgrad_kernel) with four converter warps -- the halo kernels keep the whole filter resident in
  // shared memory and have no room for the second copy of every tile.
  const int x3 = e->train.precision == SB_PREC_3XTF32 ? 1 : 0;
  int dev_smem = 0;
  SB_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device));
  for (size_t li = 0; li < e->L.size(); ++li) {
    LayerPlan& L = e->L[li];
    if (L.d.kind == SB_LAYER_DENSE) {
      // MatMul and its gradients on tcgen05 when the GEMM is big enough to matter and the strides are 16-byte multiples
      const int M = (int)L.in.d[0], K = (int)L.in.d[1], N = (int)L.out.d[1];
      if (N < 32 || K < 32 || M < 32 || (double)M * N * K < (double)(1 << 22)) continue;
      float* x = e->acts[L.buf_in];
      float* z = e->acts[L.buf_out];
      float* dz = e->dacts[L.buf_out];
      float* dx = e->dacts[L.buf_in];
      float* w = e->arena + e->params[L.p_w].offset;
      float* dw = e->grads + e->params[L.p_w].offset;
      DenseTcPlan* D = new DenseTcPlan();
      D->fwd = tc_gemm_create(0, 1, x, 1, w, 1, z, 1, M, N, K, 0, nullptr, 0, e->device, x3);
      D->wgrad = tc_gemm_create(1, 1, x, 1, dz, 1, dw, 1, K, N, M, 1, e->ws, e->ws_floats, e->device, x3);
      const bool want_dgrad = L.need_dx && dx != nullptr;
      if (want_dgrad) D->dgrad = tc_gemm_create(0, 0, dz, 1, w, 1, dx, 1, M, K, N, 1, e->ws, e->ws_floats, e->device, x3);
      if (!D->fwd || !D->wgrad || (want_dgrad && !D->dgrad)) {
        tc_gemm_destroy(D->fwd); tc_gemm_destroy(D->wgrad); tc_gemm_destroy(D->dgrad);
        delete D;
        continue;
      }
      // fold the following in-place Bias and (non-softmax) Activate into the forward epilogue
      const float* bias = nullptr;
      int act = 0;
      float thr = 0.f;
      size_t nx = li + 1;
      if (nx < e->L.size() && e->L[nx].d.kind == SB_LAYER_BIAS && e->L[nx].inplace) {
        bias = e->arena + e->params[e->L[nx].p_w].offset;
        e->L[nx].skip_fwd = true;
        ++nx;
      }
      if (nx < e->L.size() && e->L[nx].d.kind == SB_LAYER_ACTIVATE && e->L[nx].inplace && e->L[nx].d.activation != SB_ACT_SOFTMAX &&
          e->L[nx].d.activation != SB_ACT_IDENTITY && (nx == li + 1 || bias)) {
        act = e->L[nx].d.activation;
        thr = e->L[nx].d.relu_threshold;
        e->L[nx].skip_fwd = true;
      }
      tc_gemm_set_epilogue(D->fwd, bias, act, thr);
      L.tc_kind = 2;
      L.tc_plan = D;
      continue;
    }
    if (L.d.kind != SB_LAYER_CONV2D || !conv_fits(L.cg)) continue;
    const ConvGeom& g = L.cg;
    ConvTcPlan* P = new ConvTcPlan();
    float* x = e->acts[L.buf_in];
    float* y = e->acts[L.buf_out];
    float* dy = e->dacts[L.buf_out];
    float* dx = e->dacts[L.buf_in];
    float* w = e->arena + e->params[L.p_w].offset;
    const int taps = g.KH * g.KW;
    // a following in-place ReLU is applied in the fprop epilogue
    if (li + 1 < e->L.size() && e->L[li + 1].d.kind == SB_LAYER_ACTIVATE && e->L[li + 1].d.activation == SB_ACT_RELU &&
        e->L[li + 1].inplace) {
      P->relu_fused = true;
    }
    // ---- fprop: tile = BH rows of the OH x OW output grid ----
    {
      IgemmParams& p = P->fwd;
      p.Hg = g.OH; p.Wg = g.OW; p.BH = std::max(1, std::min(128 / g.OW, g.OH));
      p.tiles_per_img = (g.OH + p.BH - 1) / p.BH;
      p.n_tiles = g.N * p.tiles_per_img;
      igemm_small_grid(p, g.N);
      p.KH = g.KH; p.KW = g.KW; p.kchunks = g.Cin / 32; p.sign = +1; p.N = g.Cout; p.b_mn_major = 1;
      p.nsplit = igemm_pick_nsplit(p.n_tiles, taps * p.kchunks, p.BH * p.Wg * 128 * p.ipt, g.Cout);
      p.N = g.Cout / p.nsplit;
      p.b_rows_per_tap = g.Cin;
      p.relu = P->relu_fused ? 1 : 0;
      p.thr = P->relu_fused ? e->L[li + 1].d.relu_threshold : 0.f;
      p.out = y;
      p.ldc = g.Cout; p.n0 = 0; p.pt = g.PT; p.pl = g.PL; p.x3 = x3;
      p.stages = std::max(2, std::min(8, (dev_smem - 2048) / ((x3 ? 2 : 1) * (kATileBytes + p.N * 128))));
      uint64_t xd[4] = {(uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N};
      uint32_t xb[4] = {32, (uint32_t)p.Wg, (uint32_t)p.BH, (uint32_t)p.ipt};
      P->map_x_fwd = make_map(x, 4, xd, xb);
      uint64_t wd[2] = {(uint64_t)g.Cout, (uint64_t)taps * g.Cin};
      uint32_t wb[2] = {32, 32};
      P->map_w_fwd = make_map(w, 2, wd, wb, true);
      P->fwd_smem = igemm_smem(p);
      P->fwd_grid = std::min(p.n_tiles * p.nsplit, kNumSMs);
      const char* h = getenv("SB_TC_HALO");
      const bool halo_on = !(h && h[0] == '0') && !x3;
      // the whole filter resident, or -- when it does not fit shared memory (cfg5 conv2: 295 KB) -- output-channel slices of >= 64
      // columns, each CTA keeping its slice resident and walking every tile (SB_TC_HALO_SLICES=0: the per-tap kernel for those)
      const char* hs = getenv("SB_TC_HALO_SLICES");
      const int max_slices = (hs && hs[0] == '0') ? 1 : 4;
      for (int ns = 1; halo_on && ns <= max_slices && !P->halo_fwd; ns *= 2) {
        if (g.Cout % (32 * ns) || (ns > 1 && g.Cout / ns < 64)) break;
        if (plan_halo(P->hfwd, &P->map_x_h, &P->map_y_out, x, g.Cin, g.W, g.H, g.N, g.OH, g.OW, g.KH, g.KW, g.Cout, 1, g.Cin, 0, -g.PL, -g.PT, y,
                      dev_smem, &P->hfwd_smem, 0, ns)) {
          P->hfwd.relu = p.relu; P->hfwd.thr = p.thr;
          P->hfwd_slices = ns;
          P->hfwd_grid = std::max(1, std::min(P->hfwd.n_tiles, kNumSMs / ns));
          P->halo_fwd = true;
        }
      }
    }
    // ---- dgrad: tile = BH rows of the H x W input grid; source dY shifted by -tap with zero fill ----
    P->has_dgrad = L.need_dx && dx != nullptr;
    if (P->has_dgrad) {
      IgemmParams& p = P->dg;
      p.Hg = g.H; p.Wg = g.W; p.BH = std::max(1, std::min(128 / g.W, g.H));
      p.tiles_per_img = (g.H + p.BH - 1) / p.BH;
      p.n_tiles = g.N * p.tiles_per_img;
      igemm_small_grid(p, g.N);
      p.nsplit = 1;
      p.KH = g.KH; p.KW = g.KW; p.kchunks = g.Cout / 32; p.sign = -1; p.N = g.Cin; p.b_mn_major = 0;
      p.b_rows_per_tap = g.Cin;
      p.relu = 0; p.thr = 0.f;
      p.out = dx;
      p.ldc = g.Cin; p.n0 = 0; p.pt = -g.PT; p.pl = -g.PL; p.x3 = x3;
      p.stages = std::max(2, std::min(8, (dev_smem - 2048) / ((x3 ? 2 : 1) * (kATileBytes + p.N * 128))));
      uint64_t yd[4] = {(uint64_t)g.Cout, (uint64_t)g.OW, (uint64_t)g.OH, (uint64_t)g.N};
      uint32_t yb[4] = {32, (uint32_t)p.Wg, (uint32_t)p.BH, (uint32_t)p.ipt};
      P->map_dy_dg = make_map(dy, 4, yd, yb);
      uint64_t wd[2] = {(uint64_t)g.Cout, (uint64_t)taps * g.Cin};
      uint32_t wb[2] = {32, (uint32_t)g.Cin};
      P->map_w_dg = make_map(w, 2, wd, wb);
      P->dg_smem = igemm_smem(p);
      P->dg_grid = std::min(p.n_tiles, kNumSMs);
      const char* h = getenv("SB_TC_HALO");
      const bool halo_on = !(h && h[0] == '0') && !x3;
      if (halo_on && plan_halo(P->hdg, &P->map_dy_h, &P->map_dx_out, dy, g.Cout, g.OW, g.OH, g.N, g.H, g.W, g.KH, g.KW, g.Cin, 0, g.Cin, 1,
                               -(g.KW - 1 - g.PL), -(g.KH - 1 - g.PT), dx, dev_smem, &P->hdg_smem)) {
        P->hdg_grid = std::min(P->hdg.n_tiles, kNumSMs);
        P->halo_dg = true;
      }
    }
    // ---- wgrad: one launch per (group of taps whose accumulators fit 512 TMEM columns) x (64/128 output channels) ----
    P->wg_ok = g.Cout % 64 == 0;
    if (P->wg_ok) {
      const int Mh = g.Cout % 128 == 0 ? 128 : 64;
      const int tpp = std::max(1, std::min(taps, 512 / g.Cin));  // taps per launch
      WgradParams p{};
      p.KH = g.KH; p.KW = g.KW; p.Cin = g.Cin; p.Cout = g.Cout; p.Mh = Mh; p.pt = g.PT; p.pl = g.PL; p.x3 = x3;
      p.WK = (g.OW + 7) / 8 * 8;
      const int box_row_bytes = p.WK * 128;  // one grid row of one 32-channel box
      const int per_bh = (x3 ? 2 : 1) * (Mh / 32 + tpp * (g.Cin / 32)) * box_row_bytes;
      // deepest chunk that still leaves >= 2 pipeline stages
      p.BH = std::max(1, std::min(g.OH, (dev_smem - 4096) / 2 / per_bh));
      while (p.BH > 1 && 256 / p.WK < p.BH) --p.BH;  // TMA box dims <= 256
      p.chunks_per_img = (g.OH + p.BH - 1) / p.BH;
      p.n_chunks = g.N * p.chunks_per_img;
      const int n_slices = ((taps + tpp - 1) / tpp) * (g.Cout / Mh);
      const int full_grid = std::min(p.n_chunks, kNumSMs);  // what the halo kernel below would use
      P->wg_grid = n_slices <= kWgradMaxSlices ? std::max(1, std::min(p.n_chunks, kNumSMs / n_slices)) : full_grid;  // the K split
      (void)full_grid;  // the partial buffer is allocated below, once the halo form's grid is known too
      p.nt_max = tpp;
      {
        const int stage_bytes = (x3 ? 2 : 1) * (((p.BH * (Mh / 32 + tpp * (g.Cin / 32)) * box_row_bytes + 1023) / 1024) * 1024);
        p.stages = std::max(1, std::min(4, (dev_smem - 4096) / stage_bytes));
        P->wg_smem = p.stages * stage_bytes + 16 * (3 * p.stages + 2) + 16 + 1024;
        if (P->wg_smem > dev_smem) P->wg_ok = false;
      }
      // every (tap group, channel half) slice in ONE launch (gridDim.y); more than 16 slices: one launch per 16
      WgradParams q = p;
      for (int t0 = 0; t0 < taps; t0 += tpp)
        for (int co0 = 0; co0 < g.Cout; co0 += Mh) {
          q.s_t0[q.n_slices] = (unsigned char)t0;
          q.s_nt[q.n_slices] = (unsigned char)std::min(tpp, taps - t0);
          q.s_co0[q.n_slices] = (unsigned short)co0;
          if (++q.n_slices == kWgradMaxSlices) { P->wg_list.push_back(q); q = p; }
        }
      if (q.n_slices) P->wg_list.push_back(q);
      {
        const char* dbg = getenv("SB_TC_DEBUG");
        if (dbg && dbg[0] == '1') {
          long long* d = nullptr;
          SB_CUDA(cudaMalloc((void**)&d, (size_t)kNumSMs * 64 * sizeof(long long)));
          SB_CUDA(cudaMemset(d, 0, (size_t)kNumSMs * 64 * sizeof(long long)));
          for (WgradParams& q : P->wg_list) q.dbg = d;
          P->hwg.dbg = d;
        }
      }
      P->wg = P->wg_list[0];
      uint64_t yd[4] = {(uint64_t)g.Cout, (uint64_t)g.OW, (uint64_t)g.OH, (uint64_t)g.N};
      uint32_t yb[4] = {32, (uint32_t)p.WK, (uint32_t)p.BH, 1};
      P->map_dy_wg = make_map(dy, 4, yd, yb, true);
      uint64_t xd[4] = {(uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N};
      P->map_x_wg = make_map(x, 4, xd, yb, true);
    }
    // ---- wgrad, halo form: KW <= 4 taps share one M = 128 MMA; all KH * Cin/32 accumulators must fit TMEM ----
    {
      // Default since the issue loop keeps its TMEM addresses in uniform registers and the epilogue stores whole lines: one X
      // box per chunk instead of one per tap removes the per-tap kernel's 9x redundant L2->SMEM traffic, which is what bounds it
      // (profiles/r1_notes.md item 8).  SB_TC_HALO_WGRAD=0 selects the per-tap kernel.
      const char* h = getenv("SB_TC_HALO_WGRAD");
      const bool on = !(h && h[0] == '0') && !x3;
      WgradHaloParams& p = P->hwg;
      const int n_acc = g.KH * (g.Cin / 32);
      // output-channel slices (gridDim.y): the widest Ns (multiple of 32 dividing Cout) whose accumulators fit TMEM; below 64 the MMA
      // is bound by its operand fetch and X would be re-read once per slice, so the per-tap kernel keeps those shapes
      int Ns = 0;
      for (int c = g.Cout; c >= 32; c -= 32)
        if (g.Cout % c == 0 && n_acc * c <= 512) { Ns = c; break; }
      const int n_slices = Ns ? g.Cout / Ns : 0;
      const int slice_ctas = Ns ? std::max(1, kNumSMs / n_slices) : 0;
      if (on && P->wg_ok && g.KW <= 4 && Ns && (Ns == g.Cout || Ns >= 64)) {
        // pitch P >= OW + KW - 1 and rows per chunk BH: BH*P % 8 == 0, boxes <= 256 rows; minimise rounds x k-steps per chunk
        double best = 1e30;
        int bP = 0, bBH = 0;
        for (int Pc = g.OW + g.KW - 1; Pc <= g.OW + g.KW - 1 + 8 && Pc <= 256; ++Pc)
          for (int BH = 1; BH <= g.OH; ++BH) {
            if ((BH * Pc) % 8 || (BH + g.KH - 1) > 256) continue;
            const int rows_x = (BH + g.KH - 1) * Pc + g.KW - 1;
            const int stage = (g.Cin / 32) * ((rows_x + 7) / 8) * 1024 + (Ns / 32) * BH * Pc * 128;
            if (2 * stage + 4096 > dev_smem) continue;
            if (4 * 32 * (Ns + 4) * 4 > 2 * stage) continue;  // the epilogue's transpose scratch reuses the idle stages
            const long long chunks = (long long)g.N * ((g.OH + BH - 1) / BH);
            // cycles per CTA: rounds x max(MMA issue, TMA fill at ~25 B/cycle/SM with every SM pulling) + the first fill
            // (MMA rates and the fill rate: profiles/r1_notes.md items 8)
            const double mma = Ns <= 32 ? 40 : Ns <= 64 ? 48 : Ns / 2;
            const double fill = stage / 25.0;
            const double cost = (double)((chunks + slice_ctas - 1) / slice_ctas) * std::max((BH * Pc / 8) * n_acc * mma + 150.0, fill) + fill;
            if (cost < best) { best = cost; bP = Pc; bBH = BH; }
          }
        if (bP) {
          p.BH = bBH; p.P = bP; p.KH = g.KH; p.KW = g.KW; p.Cin = g.Cin; p.Cout = g.Cout; p.OH = g.OH; p.Ns = Ns;
          p.ox = -g.PL; p.oy = -g.PT;
          p.chunks_per_img = (g.OH + p.BH - 1) / p.BH;
          p.n_chunks = g.N * p.chunks_per_img;
          p.dy_box_bytes = p.BH * p.P * 128;
          p.dy_tile_bytes = p.dy_box_bytes;  // BH*P % 8 == 0 => multiple of 1024
          p.x_box_bytes = (p.BH + g.KH - 1) * p.P * 128;
          p.x_tile_bytes = (((p.BH + g.KH - 1) * p.P + g.KW - 1 + 7) / 8) * 1024;
          const int stage_bytes = (g.Cin / 32) * p.x_tile_bytes + (Ns / 32) * p.dy_tile_bytes;
          p.stages = std::max(2, std::min(6, (dev_smem - 4096) / stage_bytes));
          P->hwg_grid = (int)std::min<long long>(p.n_chunks, slice_ctas);
          P->hwg_slices = n_slices;
          P->hwg_smem = p.stages * stage_bytes + 16 * (2 * p.stages + 2) + 16 + 1024;
          if (P->hwg_smem <= dev_smem) {  // the partial buffer is sized for min(n_chunks, SMs) CTAs
            uint64_t yd[4] = {(uint64_t)g.Cout, (uint64_t)g.OW, (uint64_t)g.OH, (uint64_t)g.N};
            uint32_t yb[4] = {32, (uint32_t)p.P, (uint32_t)p.BH, 1};
            P->map_dy_hwg = make_map(dy, 4, yd, yb, true);
            uint64_t xd[4] = {(uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N};
            uint32_t xb[4] = {32, (uint32_t)p.P, (uint32_t)(p.BH + g.KH - 1), 1};
            P->map_x_hwg = make_map(x, 4, xd, xb, true);
            P->halo_wg = true;
            { const char* ex = getenv("SB_WG_EXPERIMENT"); p.dbg_mode = ex ? atoi(ex) : 0; }
            if (p.dbg) fprintf(stderr, "[sb] halo wgrad: BH %d P %d chunks %d stages %d stage_bytes %d grid %d\n", p.BH, p.P, p.n_chunks, p.stages, stage_bytes, P->hwg_grid);
          }
        }
      }
    }
    if (P->fwd_smem > dev_smem || P->dg_smem > dev_smem) {
      delete P;
      continue;
    }
    if (P->wg_ok) {
      // one partial filter gradient per K split: the per-tap kernel writes wg_grid of them, the halo form hwg_grid (its chunks are cut
      // differently, so it can have MORE of them than the per-tap kernel)
      const int parts = std::max(P->wg_grid, P->halo_wg ? P->hwg_grid : 0);
      SB_CUDA(cudaMalloc((void**)&P->wg_partial, (size_t)parts * taps * g.Cin * g.Cout * sizeof(float)));
      for (WgradParams& q : P->wg_list) q.partial = P->wg_partial;
      P->wg.partial = P->wg_partial;
      P->hwg.partial = P->wg_partial;
    }
    L.tc_kind = 1;
    L.tc_plan = P;
    if (P->relu_fused) e->L[li + 1].skip_fwd = true;  // forward only; the backward of ReLU still runs
  }
  static bool attr_set = false;
  if (!attr_set) {
    SB_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    SB_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    SB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    SB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<3, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    SB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<3, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    SB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<3, 3, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    SB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<3, 3, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    SB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<0, 0, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    SB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<3, 3, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    SB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<3, 3, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    SB_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    SB_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    SB_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
    attr_set = true;
  }
}

static void head_tc_free(LayerPlan& L);  // defined with the kernel below
static void head_bwd_tc_free(LayerPlan& L);
void tc_free_layers(sb_engine* e) {
  for (auto& L : e->L) {
    head_tc_free(L);
    head_bwd_tc_free(L);
    if (L.tc_plan && L.tc_kind == 2) {
      DenseTcPlan* D = (DenseTcPlan*)L.tc_plan;
      tc_gemm_destroy(D->fwd); tc_gemm_destroy(D->wgrad); tc_gemm_destroy(D->dgrad);
      delete D;
      L.tc_plan = nullptr;
      continue;
    }
    if (L.tc_plan) {
      ConvTcPlan* P = (ConvTcPlan*)L.tc_plan;
      for (HaloParams* hp : {&P->hfwd, &P->hdg}) {
        if (!hp->dbg) continue;  // developer timeline of the LAST launch, one line per CTA
        std::vector<long long> h((size_t)kNumSMs * 64);
        cudaMemcpy(h.data(), hp->dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        FILE* f = fopen(hp == &P->hfwd ? "gpurun_out/halo_fwd_timeline.txt" : "gpurun_out/halo_dgrad_timeline.txt", "w");
        if (f) {
          for (int c = 0; c < kNumSMs; ++c) {
            for (int j = 0; j < 48; ++j) fprintf(f, "%lld ", h[(size_t)c * 64 + j] ? h[(size_t)c * 64 + j] - h[(size_t)c * 64] : -1);
            fprintf(f, "\n");
          }
          fclose(f);
        }
        cudaFree(hp->dbg);
      }
      if (!P->wg_list.empty() && P->wg_list[0].dbg) {
        std::vector<long long> h((size_t)kNumSMs * 64);
        cudaMemcpy(h.data(), P->wg_list[0].dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        FILE* f = fopen("gpurun_out/wgrad_timeline.txt", "w");
        if (f) {
          for (int c = 0; c < kNumSMs; ++c) {
            for (int j = 0; j < 34; ++j) fprintf(f, "%lld ", h[(size_t)c * 64 + j] ? h[(size_t)c * 64 + j] - h[(size_t)c * 64] : -1);
            fprintf(f, "\n");
          }
          fclose(f);
        }
        cudaFree(P->wg_list[0].dbg);
      }
      cudaFree(P->wg_partial);
      delete P;
      L.tc_plan = nullptr;
    }
  }
}

void tc_params_changed(sb_engine*) {}  // the kernels read the filter straight from the arena

// Called by the fusion planner (engine.cu) once it knows that nothing but the following Pool2D(2x2, stride 1, VALID) reads this
// conv's activations: re-plans the halo fprop with the pooled epilogue.  false: not available (no halo fprop, geometry, shared memory)
bool tc_conv_enable_pool_fusion(sb_engine* e, LayerPlan& L, float* pooled, unsigned char* map) {
  ConvTcPlan* P = (ConvTcPlan*)L.tc_plan;
  // OFF by default (SB_TC_POOL_FUSION=1 turns it on; parity-tested bit-identical either way).  Measured on cfg2: the pooled
  // epilogue's shared-memory traffic (32 KB of staging writes + 90 KB of window reads per tile) competes with the MMAs' own
  // operand fetch, which already saturates shared memory at N = 64 (6 KB per MMA at 128 B/cycle = the 48 cycles an MMA takes): a
  // tile goes from ~1900 to ~4500 cycles whether 8 or 24 warps pool (SB_TC_DEBUG timeline), fprop 15 -> 23 us, step 1035 k -> 971 k
  // samples/s -- more than the deleted pool kernel (5 us in the step) gives back.
  const char* on = getenv("SB_TC_POOL_FUSION");
  if (!P || L.tc_kind != 1 || !P->halo_fwd || P->hfwd_slices != 1 || !(on && on[0] == '1')) return false;
  const ConvGeom& g = L.cg;
  int dev_smem = 0;
  SB_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device));
  HaloParams hp = P->hfwd;
  CUtensorMap mx, mo;
  int smem = 0;
  if (!plan_halo(hp, &mx, &mo, e->acts[L.buf_in], g.Cin, g.W, g.H, g.N, g.OH, g.OW, g.KH, g.KW, g.Cout, 1, g.Cin, 0, -g.PL, -g.PT,
                 e->acts[L.buf_out], dev_smem, &smem, 1))
    return false;
  hp.relu = P->hfwd.relu; hp.thr = P->hfwd.thr; hp.dbg = P->hfwd.dbg;
  hp.pooled = pooled;
  hp.pmap = reinterpret_cast<unsigned*>(map);
  P->hfwd = hp;
  P->map_x_h = mx;
  P->map_y_out = mo;
  P->hfwd_smem = smem;
  P->hfwd_grid = std::min(hp.n_tiles, kNumSMs);
  return true;
}

bool tc_dense_fwd(sb_engine* e, LayerPlan& L, int) {
  if (L.tc_kind != 2) return false;
  tc_gemm_run(((DenseTcPlan*)L.tc_plan)->fwd, 0, 0, 0, e->stream);
  return true;
}
bool tc_dense_wgrad(sb_engine* e, LayerPlan& L, int) {
  if (L.tc_kind != 2) return false;
  tc_gemm_run(((DenseTcPlan*)L.tc_plan)->wgrad, 0, 0, 0, e->stream);
  return true;
}
bool tc_dense_dgrad(sb_engine* e, LayerPlan& L, int) {
  if (L.tc_kind != 2 || !((DenseTcPlan*)L.tc_plan)->dgrad) return false;
  tc_gemm_run(((DenseTcPlan*)L.tc_plan)->dgrad, 0, 0, 0, e->stream);
  return true;
}

bool tc_conv_fwd(sb_engine* e, LayerPlan& L, int) {
  ConvTcPlan* P = (ConvTcPlan*)L.tc_plan;
  if (!P || L.tc_kind != 1) return false;
  if (P->halo_fwd) {
    const bool k331 = P->hfwd.KH == 3 && P->hfwd.KW == 3 && P->hfwd.kchunks == 1, k332 = P->hfwd.KH == 3 && P->hfwd.KW == 3 && P->hfwd.kchunks == 2;
    if (P->hfwd.pool) {
      auto kern = k331 ? conv_halo_kernel<3, 3, 1, true> : k332 ? conv_halo_kernel<3, 3, 2, true> : conv_halo_kernel<0, 0, 0, true>;
      launch_k(kern, P->hfwd_grid, kHaloPoolThreads, P->hfwd_smem, e->stream, P->map_x_h, P->map_w_fwd, P->map_y_out, P->hfwd);
    } else if (P->hfwd.fold && (k331 || k332)) {
      auto kern = k331 ? conv_halo_kernel<3, 3, 1, false, true> : conv_halo_kernel<3, 3, 2, false, true>;
      launch_k(kern, dim3(P->hfwd_grid, P->hfwd_slices), kIgemmThreadsX3, P->hfwd_smem, e->stream, P->map_x_h, P->map_w_fwd, P->map_y_out, P->hfwd);
    } else {
      auto kern = k331 ? conv_halo_kernel<3, 3, 1> : k332 ? conv_halo_kernel<3, 3, 2> : conv_halo_kernel<0, 0, 0>;
      launch_k(kern, dim3(P->hfwd_grid, P->hfwd_slices), kIgemmThreads, P->hfwd_smem, e->stream, P->map_x_h, P->map_w_fwd, P->map_y_out, P->hfwd);
    }
    SB_LAUNCHED();
    return true;
  }
  launch_k(conv_igemm_kernel, P->fwd_grid, P->fwd.x3 ? kIgemmThreadsX3 : kIgemmThreads, P->fwd_smem, e->stream, P->map_x_fwd, P->map_w_fwd, P->fwd);
  SB_LAUNCHED();
  return true;
}

// =====================================================================================================================
// Skinny classifier head forward on the tensor cores:  Z[M <= 128, N <= 16] = X[M, K] . W[K, N], split-K.
// CTA c owns K range [c*KC, c*KC + KC): its X tiles [128 rows x 32 k] arrive by TMA (rows >= M are zero fill), one mbarrier per
// tile so the MMAs start with the first tile; the W slab is tiny (KC x N floats, row pitch N*4 bytes is not TMA-able) and is
// written by the threads straight into the K-major SWIZZLE_128B layout TMA would have produced (row n, 16-byte chunk
// (k/4) ^ (n%8)).  D[128, 16] accumulates in TMEM; partial[c][m][n] feeds the fused bias+softmax+loss kernel as before.
// The FFMA version (kernels_fast.cu) spends as long on its multiply phase as on the fill; here only the fill remains.
// =====================================================================================================================
struct HeadFwdParams {
  int M, K, N, KC, chunks;
  int prewait;  // common.cuh prewait_ok(): the W slab may be read before the dependency wait
  const float* w;
  float* partial;
};
constexpr int kHeadTcThreads = 192;
constexpr int kHeadTcMaxTiles = 10;

__global__ void __launch_bounds__(kHeadTcThreads, 1) head_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_x, const HeadFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nkt_max = p.KC / 32;
  uint8_t* sm_a = smem;                                  // nkt x [128 rows x 128 B]
  uint8_t* sm_b = smem + (size_t)nkt_max * 16384;        // nkt x [16 rows x 128 B]
  uint64_t* full_bar = (uint64_t*)(sm_b + (size_t)nkt_max * 2048);
  uint64_t* done_bar = full_bar + kHeadTcMaxTiles;
  uint32_t* tmem_slot = (uint32_t*)(done_bar + 1);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * p.KC;
  const int nkt = min(nkt_max, (p.K - k0) / 32);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x);
    for (int t = 0; t < nkt_max; ++t) mbar_init(&full_bar[t], 1);
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // W slab: element (n, k) of tile t goes to row n (128 B), chunk ((k%32)/4) ^ (n%8), word k%4.  When the kernel launched just
  // before this one does not write weights (p.prewait, common.cuh) it is staged before the programmatic-dependency wait
  if (!p.prewait) pdl_wait();
  for (int i = threadIdx.x; i < nkt * 32 * 16; i += blockDim.x) {
    const int n = i & 15, k = i >> 4;
    const float v = n < p.N ? __ldg(p.w + (size_t)(k0 + k) * p.N + n) : 0.f;
    const int t = k >> 5, kk = k & 31;
    *reinterpret_cast<float*>(sm_b + (size_t)t * 2048 + n * 128 + ((((kk >> 2) ^ (n & 7)) << 4) | ((kk & 3) << 2))) = v;
  }
  fence_proxy_async();
  if (p.prewait) pdl_wait();
  pdl_trigger();  // dependents may start their own on-chip setup now (they still wait for this grid before touching its output)
  __syncthreads();
  if (warp == 0) {
    if (lane == 0)
      for (int t = 0; t < nkt; ++t) {
        mbar_arrive_expect_tx(&full_bar[t], 16384u);
        tma_load_2d(sm_a + (size_t)t * 16384, &map_x, &full_bar[t], k0 + t * 32, 0);
      }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_tf32(128, 16, 0, 0);
    const uint64_t d = smem_desc_sw128(0, 16, 1024);
    const uint32_t d_hi = (uint32_t)(d >> 32), d_lo = (uint32_t)d;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a0 = d_lo + (smem_u32(sm_a) >> 4), b0 = d_lo + (smem_u32(sm_b) >> 4);
    uint32_t accumulate = 0;
    for (int t = 0; t < nkt; ++t) {
      mbar_wait(&full_bar[t], 0);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        umma_tf32_elect(tmem_u, ((uint64_t)d_hi << 32) | (a0 + t * 1024 + ks * 2), ((uint64_t)d_hi << 32) | (b0 + t * 128 + ks * 2), idesc,
                        accumulate);
        accumulate = 1;
      }
    }
    umma_commit_elect(done_bar);
  } else {
    const int sub = warp & 3, row = sub * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    float v[32];
    tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16), v);
    tmem_ld_wait();
    if (row < p.M) {
      float* o = p.partial + ((size_t)blockIdx.x * p.M + row) * p.N;
#pragma unroll
      for (int n = 0; n < 16; ++n)
        if (n < p.N) o[n] = v[n];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

struct HeadTcPlan {
  CUtensorMap map_x;
  HeadFwdParams p;
  const float* x;
  int smem;
};

static void head_tc_free(LayerPlan& L) {
  delete (HeadTcPlan*)L.head_tc;
  L.head_tc = nullptr;
}

bool tc_head_fwd(sb_engine* e, LayerPlan& L, const float* x, const float* w, float* partial, size_t partial_floats, int* chunks_out) {
  static const bool enabled = [] { const char* v = getenv("SB_TC_HEAD"); return !(v && v[0] == '0'); }();
  if (!enabled || e->train.precision != SB_PREC_TF32) return false;
  const int M = (int)L.in.d[0], K = (int)L.in.d[1], N = (int)L.out.d[1];
  if (M > 128 || N > 16 || (K % 32) || K < 32 * kNumSMs || (((uintptr_t)x) & 127)) return false;
  HeadTcPlan* P = (HeadTcPlan*)L.head_tc;
  if (!P || P->x != x) {
    delete P;
    P = new HeadTcPlan();
    const int tiles = K / 32;
    int per = (tiles + kNumSMs - 1) / kNumSMs;
    if (per > kHeadTcMaxTiles) { delete P; L.head_tc = nullptr; return false; }
    P->p.M = M; P->p.K = K; P->p.N = N; P->p.KC = per * 32; P->p.chunks = (tiles + per - 1) / per;
    uint64_t xd[2] = {(uint64_t)K, (uint64_t)M};
    uint32_t xb[2] = {32, 128};
    P->map_x = make_map(x, 2, xd, xb);
    P->x = x;
    P->smem = per * (16384 + 2048) + 16 * (kHeadTcMaxTiles + 2) + 1024;
    SB_CUDA(cudaFuncSetAttribute(head_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P->smem));
    L.head_tc = P;
  }
  if ((size_t)P->p.chunks * M * N > partial_floats) return false;
  P->p.w = w;
  P->p.partial = partial;
  P->p.prewait = prewait_ok();
  launch_k(head_fwd_tc_kernel, P->p.chunks, kHeadTcThreads, P->smem, e->stream, P->map_x, P->p);
  SB_LAUNCHED();
  *chunks_out = P->p.chunks;
  return true;
}

// =====================================================================================================================
// Skinny classifier head BACKWARD on the tensor cores (TF32 mode, M <= 128 rows, N <= 16 outputs, K % 128 == 0):
//     dW[k][n] = sum_m X[m][k] G[m][n]          dX[m][k] = sum_n G[m][n] W[k][n]
// CTA c owns the 128 k columns [128c, 128c + 128).  The FFMA kernel (kernels_fast.cu dense_head_bwd_kernel) issues ~40 instructions
// per (row, two columns) on 13 warps per SM (18.8 us on cfg2: issue-bound at 20 % occupancy); here
//   * dW tile D1[M = 128 k, N = 16] : A = the X tile straight from memory, MN-major (k contiguous), four TMA boxes [32 k x MP rows]
//     (MP = M rounded up to 8; rows >= M are TMA zero fill), K dimension = the batch rows; B = G^T written by the threads in the
//     K-major SWIZZLE_128B layout (row n, 16-byte chunk ((m % 32) / 4) ^ (n % 8));  MP/8 MMAs.
//   * dX tile D2[M = 128 rows, N = 128 k] : A = G, B = the W slab [128 k x 16 n], both K-major and written by the threads
//     (row pitches of N * 4 bytes are not TMA-able);  two MMAs -- issued first, they do not wait for the X tile.
//   X is read once by TMA, dX written once as 512-byte row pieces, dW as one contiguous slab: the kernel is bound by those bytes.
// =====================================================================================================================
struct HeadBwdParams {
  int M, MP, K, N;
  int prewait;
  const float* g;   // [M, N]
  const float* w;   // [K, N]
  float* dw;        // [K, N]
  float* dx;        // [M, K] or null
};
constexpr int kHeadBwdThreads = 192;

__global__ void __launch_bounds__(kHeadBwdThreads, 2) head_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dx,
                                                                         const HeadBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int box_bytes = p.MP * 128;                       // one [32 k x MP rows] box of X
  const int a_bytes = ((4 * box_bytes + 1023) / 1024) * 1024;
  uint8_t* sm_x = smem;                                   // 4 boxes, MN-major (SWIZZLE_128B_ATOM_32B)
  uint8_t* sm_gt = sm_x + a_bytes;                        // G^T : 4 tiles of [16 rows n x 128 B (32 m)]
  uint8_t* sm_g = sm_gt + 4 * 2048;                       // G   : [128 rows m x 128 B (n)]
  uint8_t* sm_w = sm_g + 16384;                           // W   : [128 rows k x 128 B (n)]
  // barriers behind the operands AND behind the 64 KB the dX staging reuses once the MMAs are done
  uint64_t* x_bar = (uint64_t*)(smem + max(a_bytes + 4 * 2048 + 2 * 16384, 4 * 16384));
  uint64_t* done_bar = x_bar + 1;
  uint32_t* tmem_slot = (uint32_t*)(done_bar + 1);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_x);
    mbar_init(x_bar, 1);
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {  // two allocations (128 + 32 columns, not one of 256): two CTAs per SM leave TMEM for an overlapping dependent kernel
    tmem_alloc_keep_permit(tmem_slot, 128);  // dX tile
    tmem_alloc(tmem_slot + 1, 32);   // dW tile
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot[0], tmem_dw = tmem_slot[1];
  // Staging of the thread-written operands.  The sources are CONTIGUOUS in memory (128 rows of W, M rows of G, N floats each), so
  // the threads walk them linearly with every load in flight (fixed trip counts, unrolled) and scatter into the swizzled layouts;
  // the padding (n >= N, m >= M) comes from zero-filling the three areas first.
  const unsigned div_n = (65536u + p.N - 1) / p.N;  // i / N == (i * div_n) >> 16 for i < 128 * 16
  {
    float4* z = reinterpret_cast<float4*>(sm_gt);
    for (int i = threadIdx.x; i < (4 * 2048 + 2 * 16384) / 16; i += blockDim.x) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  auto stage_w = [&]() {  // element (k, n) -> row k, 16-byte chunk (n / 4) ^ (k % 8), word n % 4
    const int total = 128 * p.N;
    float v[11];
#pragma unroll
    for (int it = 0; it < 11; ++it) {
      const int i = threadIdx.x + it * kHeadBwdThreads;
      v[it] = i < total ? __ldg(p.w + (size_t)k0 * p.N + i) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < 11; ++it) {
      const int i = threadIdx.x + it * kHeadBwdThreads;
      if (i < total) {
        const int k = (int)(((unsigned)i * div_n) >> 16), n = i - k * p.N;
        *reinterpret_cast<float*>(sm_w + k * 128 + ((((n >> 2) ^ (k & 7)) << 4) | ((n & 3) << 2))) = v[it];
      }
    }
  };
  // the W slab may be staged before the dependency wait when the kernel launched just before does not write weights
  if (p.prewait) stage_w();
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) {  // the X tile: rows >= M of every box are zero fill
    mbar_arrive_expect_tx(x_bar, (uint32_t)(4 * box_bytes));
    for (int j = 0; j < 4; ++j) tma_load_2d(sm_x + (size_t)j * box_bytes, &map_x, x_bar, k0 + j * 32, 0);
  }
  if (!p.prewait) stage_w();
  {
    const int total = p.M * p.N;
    float v[11];
#pragma unroll
    for (int it = 0; it < 11; ++it) {
      const int i = threadIdx.x + it * kHeadBwdThreads;
      v[it] = i < total ? __ldg(p.g + i) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < 11; ++it) {
      const int i = threadIdx.x + it * kHeadBwdThreads;
      if (i < total) {
        const int m = (int)(((unsigned)i * div_n) >> 16), n = i - m * p.N;
        *reinterpret_cast<float*>(sm_g + m * 128 + ((((n >> 2) ^ (m & 7)) << 4) | ((n & 3) << 2))) = v[it];                 // G   (row m, n along K)
        const int t = m >> 5, mm = m & 31;
        *reinterpret_cast<float*>(sm_gt + (size_t)t * 2048 + n * 128 + ((((mm >> 2) ^ (n & 7)) << 4) | ((mm & 3) << 2))) = v[it];  // G^T (row n, m along K)
      }
    }
  }
  fence_proxy_async();
  __syncthreads();
  if (warp == 1) {
    const uint64_t kd = smem_desc_sw128(0, 16, 1024);                      // K-major operands
    const uint32_t k_hi = (uint32_t)(kd >> 32), k_lo = (uint32_t)kd;
    const uint64_t xd = smem_desc(0, (uint32_t)box_bytes, 512, 1);         // MN-major X: 32-column groups one box apart
    const uint32_t x_hi = (uint32_t)(xd >> 32), x_lo = (uint32_t)xd;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0), tmem_w = __shfl_sync(0xffffffffu, tmem_dw, 0);
    const uint32_t g0 = k_lo + ((smem_u32(sm_g) >> 4) & 0x3FFFu), w0 = k_lo + ((smem_u32(sm_w) >> 4) & 0x3FFFu);
    const uint32_t gt0 = k_lo + ((smem_u32(sm_gt) >> 4) & 0x3FFFu), x0 = x_lo + ((smem_u32(sm_x) >> 4) & 0x3FFFu);
    if (p.dx) {  // dX tile: D2[128 m, 128 k] = G[128, 16] . W_tile^T -- operands staged by the threads, no wait for TMA
      const uint32_t idesc2 = idesc_tf32(128, 128, 0, 0);
      umma_tf32_elect(tmem_u, ((uint64_t)k_hi << 32) | g0, ((uint64_t)k_hi << 32) | w0, idesc2, 0u);
      umma_tf32_elect(tmem_u, ((uint64_t)k_hi << 32) | (g0 + 2), ((uint64_t)k_hi << 32) | (w0 + 2), idesc2, 1u);
    }
    mbar_wait(x_bar, 0);
    tc_fence_after();
    const uint32_t idesc1 = idesc_tf32(128, 16, 1, 0);
    const int steps = p.MP / 8;
    for (int sidx = 0; sidx < steps; ++sidx)  // 8 batch rows per MMA: A rows are 128 B apart inside a box, B tiles hold 32 rows each
      umma_tf32_elect(tmem_w, ((uint64_t)x_hi << 32) | (x0 + sidx * 64),
                      ((uint64_t)k_hi << 32) | (gt0 + (sidx >> 2) * 128 + (sidx & 3) * 2), idesc1, sidx ? 1u : 0u);
    umma_commit_elect(done_bar);
  } else if (warp >= 2) {
    const int sub = warp & 3, row = sub * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    float v[32];
    // dW: TMEM lane = k column of the tile
    tmem_ld32(tmem_dw + ((uint32_t)(sub * 32) << 16), v);
    tmem_ld_wait();
    {
      float* o = p.dw + (size_t)(k0 + row) * p.N;
#pragma unroll
      for (int n = 0; n < 16; ++n)
        if (n < p.N) o[n] = v[n];
    }
    if (p.dx) {  // dX: TMEM lane = batch row.  Per-thread stores would be 32 row-strided 16-byte pieces per thread (measured: 5.5 of the
      // kernel's 10.4 us); instead the tile goes through the swizzled layout of four [128 rows x 32 k] TMA boxes in the (now dead)
      // operand memory and leaves as four bulk stores -- rows past the batch are clipped by the tensor map
      for (int c0 = 0; c0 < 128; c0 += 32) {
        tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        uint8_t* srow = smem + (size_t)(c0 >> 5) * 16384 + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(srow + ((j ^ (row & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        for (int c = 0; c < 4; ++c) tma_store_2d(&map_dx, smem + (size_t)c * 16384, k0 + c * 32, 0);
        tma_store_commit_group();
        tma_store_wait_all_groups();  // the staging memory is read (and the global writes are done) before the CTA exits
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
    tmem_dealloc(tmem_dw, 32);
  }
}

struct HeadBwdTcPlan {
  CUtensorMap map_x, map_dx;
  HeadBwdParams p;
  const float* x;
  const float* dx;
  int smem;
};

static void head_bwd_tc_free(LayerPlan& L) {
  delete (HeadBwdTcPlan*)L.head_bwd_tc;
  L.head_bwd_tc = nullptr;
}

// false: shape / mode not covered (the caller runs dense_head_bwd)
bool tc_head_bwd(sb_engine* e, LayerPlan& L, const float* x, const float* g, const float* w, float* dw, float* dx, cudaStream_t st) {
  // Default in TF32 mode (SB_TC_HEAD_BWD=0: the FFMA kernel; read per call: the A/B test flips it inside one process).  Measured on
  // cfg2 (batch 100, K = 30976), in-situ timeline of the replayed step: 7.8 us from this kernel's start to the next kernel's start
  // against 10.5 us for the FFMA kernel; step 91.9 -> 89.5 us (1.088 M -> 1.117 M samples/s).  With per-thread dX stores it was no
  // faster than the FFMA kernel (10.4 us, 5.5 of them in the row-strided stores): profiles/r2_notes.md 14.
  const char* off = getenv("SB_TC_HEAD_BWD");
  if ((off && off[0] == '0') || e->train.precision != SB_PREC_TF32) return false;
  const int M = (int)L.in.d[0], K = (int)L.in.d[1], N = (int)L.out.d[1];
  if (M > 128 || N > 16 || (K % 128) || K < 128 * kNumSMs || (((uintptr_t)x) & 127) || (dx && (((uintptr_t)dx) & 15))) return false;
  HeadBwdTcPlan* P = (HeadBwdTcPlan*)L.head_bwd_tc;
  if (!P || P->x != x || P->dx != dx) {
    delete P;
    P = new HeadBwdTcPlan();
    P->p.M = M; P->p.MP = (M + 7) / 8 * 8; P->p.K = K; P->p.N = N;
    uint64_t xd[2] = {(uint64_t)K, (uint64_t)M};
    uint32_t xb[2] = {32, (uint32_t)P->p.MP};
    P->map_x = make_map(x, 2, xd, xb, true);
    P->x = x;
    P->dx = dx;
    uint32_t db[2] = {32, 128};
    P->map_dx = make_map(dx ? dx : x, 2, xd, db);  // dX tile stores: K-major boxes of 32 k x 128 rows (rows >= M clipped)
    // operands, and at least the 64 KB the dX staging reuses
    P->smem = std::max(((4 * P->p.MP * 128 + 1023) / 1024) * 1024 + 4 * 2048 + 2 * 16384, 4 * 16384) + 64 + 1024;
    SB_CUDA(cudaFuncSetAttribute(head_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P->smem));
    L.head_bwd_tc = P;
  }
  P->p.g = g; P->p.w = w; P->p.dw = dw; P->p.dx = dx;
  P->p.prewait = prewait_ok();
  launch_k(head_bwd_tc_kernel, K / 128, kHeadBwdThreads, P->smem, st, P->map_x, P->map_dx, P->p);
  SB_LAUNCHED();
  return true;
}

bool tc_conv_wgrad(sb_engine* e, LayerPlan& L, int, cudaStream_t side, ReduceBatch* defer) {
  ConvTcPlan* P = (ConvTcPlan*)L.tc_plan;
  cudaStream_t st = side ? side : e->stream;
  if (!P || L.tc_kind != 1 || !P->wg_ok) return false;
  const ConvGeom& g = L.cg;
  const int n = g.KH * g.KW * g.Cin * g.Cout;
  if (P->halo_wg) {
    const int cis = P->hwg.Cin / 32;
    auto kern = P->hwg.KH == 3 && cis == 1 ? conv_wgrad_halo_kernel<3, 1>
                : P->hwg.KH == 3 && cis == 2 ? conv_wgrad_halo_kernel<3, 2> : conv_wgrad_halo_kernel<0, 0>;
    launch_k(kern, dim3(P->hwg_grid, P->hwg_slices), kWgradThreads, P->hwg_smem, st, P->map_dy_hwg, P->map_x_hwg, P->hwg);
    SB_LAUNCHED();
    if (defer && defer->add(P->wg_partial, e->grads + e->params[L.p_w].offset, n, P->hwg_grid)) return true;  // summed before the optimizer
    partial_reduce(P->wg_partial, e->grads + e->params[L.p_w].offset, n, P->hwg_grid, st);
    return true;
  }
  for (size_t i = 0; i < P->wg_list.size(); ++i) {
    launch_k(conv_wgrad_kernel, dim3(P->wg_grid, P->wg_list[i].n_slices), P->wg_list[i].x3 ? kIgemmThreadsX3 : kWgradThreads, P->wg_smem, st,
             P->map_dy_wg, P->map_x_wg, P->wg_list[i]);
    SB_LAUNCHED();
  }
  if (defer && defer->add(P->wg_partial, e->grads + e->params[L.p_w].offset, n, P->wg_grid)) return true;
  partial_reduce(P->wg_partial, e->grads + e->params[L.p_w].offset, n, P->wg_grid, st);  // fixed order, 8-way parallel
  return true;
}

bool tc_conv_dgrad(sb_engine* e, LayerPlan& L, int) {
  ConvTcPlan* P = (ConvTcPlan*)L.tc_plan;
  if (!P || L.tc_kind != 1 || !P->has_dgrad) return false;
  if (P->halo_dg) {
    const bool k331 = P->hdg.KH == 3 && P->hdg.KW == 3 && P->hdg.kchunks == 1, k332 = P->hdg.KH == 3 && P->hdg.KW == 3 && P->hdg.kchunks == 2;
    if (P->hdg.fold && (k331 || k332)) {
      auto kern = k331 ? conv_halo_kernel<3, 3, 1, false, true> : conv_halo_kernel<3, 3, 2, false, true>;
      launch_k(kern, P->hdg_grid, kIgemmThreadsX3, P->hdg_smem, e->stream, P->map_dy_h, P->map_w_dg, P->map_dx_out, P->hdg);
    } else {
      auto kern = k331 ? conv_halo_kernel<3, 3, 1> : k332 ? conv_halo_kernel<3, 3, 2> : conv_halo_kernel<0, 0, 0>;
      launch_k(kern, P->hdg_grid, kIgemmThreads, P->hdg_smem, e->stream, P->map_dy_h, P->map_w_dg, P->map_dx_out, P->hdg);
    }
    SB_LAUNCHED();
    return true;
  }
  launch_k(conv_igemm_kernel, P->dg_grid, P->dg.x3 ? kIgemmThreadsX3 : kIgemmThreads, P->dg_smem, e->stream, P->map_dy_dg, P->map_w_dg, P->dg);
  SB_LAUNCHED();
  return true;
}

}  // namespace sb

namespace sb {
void trace_set_tc(unsigned long long* p) { sb_trace_set_local(p); }  // SB_TRACE timeline buffer of this translation unit
}  // namespace sb
