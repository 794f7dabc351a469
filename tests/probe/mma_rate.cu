// Developer micro-benchmark (not a test): issue rate of tcgen05.mma kind::tf32 for the operand layouts the conv wgrad kernels use.
// One CTA, shared memory zero-filled (timing does not depend on the values), R back-to-back MMAs into one accumulator, one
// commit, wait; cycles per MMA from clock64.   build + run: see the command in profiles/r1_notes.md.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../scanet3_b200/csrc/tc_sm100.cuh"
using namespace sb::tc;

struct Cfg {
  int M, N, a_mn, b_mn;
  int a_layout, a_lbo, a_sbo, a_kstep;  // descriptor fields (bytes) + byte advance of the A start address per MMA
  int b_layout, b_lbo, b_sbo, b_kstep;
  int ksteps;                            // distinct K positions cycled through
  int a_off;                             // byte offset of the A start address (alignment experiments)
  int n_acc;                             // accumulators cycled through (TMEM columns n_acc * N)
  const char* name;
};

__global__ void __launch_bounds__(192, 1) rate_kernel(Cfg c, int R, long long* out, int poll) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t done;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x * 16; i < 160 * 1024; i += blockDim.x * 16) {
    const float v = (poll & 2) ? (float)((i * 2654435761u >> 20) & 1023) * 0.001f + 0.5f : 0.f;  // mode bit 1: non-zero operands
    *reinterpret_cast<float4*>(smem + i) = make_float4(v, v + 1.f, v - 2.f, v * 3.f);
  }
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) {
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t idesc = idesc_tf32(c.M, c.N, c.a_mn, c.b_mn);
    const uint64_t ad = smem_desc(smem_u32(smem) + c.a_off, c.a_lbo, c.a_sbo, c.a_layout);
    const uint64_t bd = smem_desc(smem_u32(smem + 96 * 1024), c.b_lbo, c.b_sbo, c.b_layout);
    const uint32_t a_hi = (uint32_t)(ad >> 32), a_lo = (uint32_t)ad, b_hi = (uint32_t)(bd >> 32), b_lo = (uint32_t)bd;
    const long long t0 = clock64();
    const uint32_t as = c.a_kstep >> 4, bs = c.b_kstep >> 4;
    for (int r = 0; r < R; r += 4) {  // 4 K positions per iteration, no index arithmetic on the issue path
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_tf32_elect(tmem + (c.n_acc > 1 ? (uint32_t)((ks % c.n_acc) * c.N) : 0u), ((uint64_t)a_hi << 32) | (a_lo + ks * as),
                        ((uint64_t)b_hi << 32) | (b_lo + ks * bs), idesc, r ? 1u : 0u);
    }
    const long long t1 = clock64();
    umma_commit_elect(&done);
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  if (warp >= 2 && (poll & 1)) mbar_wait(&done, 0);  // what the epilogue warps of the real kernels do while the MMAs run
  if (warp >= 2 && (poll & 4)) {                       // mode bit 2: generic stores stream into another part of shared memory
    float4* dst = reinterpret_cast<float4*>(smem + 170 * 1024) + (threadIdx.x - 64);
    for (int it = 0; it < 4000 && !mbar_try_wait(&done, 0); ++it) dst[(it & 3) * 128] = make_float4(1.f, 2.f, 3.f, (float)it);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  // layout codes: 0 = no swizzle, 1 = SWIZZLE_128B_BASE32B (MN-major tf32), 2 = SWIZZLE_128B (K-major)
  const Cfg cfgs[] = {
      {128, 256, 0, 0, 2, 16, 1024, 32, 2, 16, 1024, 32, 4, 0, 1, "K-major sw128  M128 N256"},
      {128, 128, 0, 0, 2, 16, 1024, 32, 2, 16, 1024, 32, 4, 0, 1, "K-major sw128  M128 N128"},
      {128, 64, 0, 0, 2, 16, 1024, 32, 2, 16, 1024, 32, 4, 0, 1, "K-major sw128  M128 N64"},
      {128, 32, 0, 0, 2, 16, 1024, 32, 2, 16, 1024, 32, 4, 0, 1, "K-major sw128  M128 N32"},
      {64, 64, 0, 0, 2, 16, 1024, 32, 2, 16, 1024, 32, 4, 0, 1, "K-major sw128  M64  N64"},
      {64, 96, 0, 0, 2, 16, 1024, 32, 2, 16, 1024, 32, 4, 0, 1, "K-major sw128  M64  N96"},
      {64, 256, 0, 0, 2, 16, 1024, 32, 2, 16, 1024, 32, 4, 0, 1, "K-major sw128  M64  N256"},
      {64, 32, 1, 1, 1, 9216, 512, 1024, 1, 9216, 512, 1024, 9, 0, 1, "MN-major both  M64  N32  (per-tap wgrad today)"},
      {128, 32, 1, 1, 1, 9216, 512, 1024, 1, 9216, 512, 1024, 9, 0, 1, "MN-major both  M128 N32"},
      {128, 64, 1, 1, 1, 9216, 512, 1024, 1, 9216, 512, 1024, 9, 0, 1, "MN-major both  M128 N64  (groups 9 KB apart)"},
      {128, 64, 1, 1, 1, 128, 512, 1024, 1, 9216, 512, 1024, 9, 0, 1, "MN-major both  M128 N64  A groups 128 B apart (halo wgrad)"},
      {128, 256, 1, 1, 1, 4096, 512, 1024, 1, 4096, 512, 1024, 4, 0, 1, "MN-major both  M128 N256 (GEMM TN)"},
      {128, 64, 1, 0, 1, 9216, 512, 1024, 2, 16, 1024, 32, 4, 0, 1, "A MN-major, B K-major  M128 N64"},
      {128, 64, 0, 1, 2, 16, 1024, 32, 1, 9216, 512, 1024, 4, 0, 1, "A K-major, B MN-major  M128 N64"},
      {128, 64, 0, 0, 0, 128, 2304, 256, 0, 128, 2304, 256, 9, 0, 1, "K-major no-swizzle M128 N64 (core matrices 128 B apart along K)"},
      {128, 64, 0, 0, 0, 2304, 128, 256, 0, 2304, 128, 256, 9, 0, 1, "K-major no-swizzle M128 N64 (LBO/SBO swapped)"},
      {128, 256, 0, 0, 0, 128, 2304, 256, 0, 128, 2304, 256, 9, 0, 1, "K-major no-swizzle M128 N256"},
      {128, 64, 1, 1, 1, 128, 512, 1024, 1, 7168, 512, 1024, 9, 3584, 1, "MN-major M128 N64 halo wgrad, A start +3584 B (not 1024-aligned)"},
      {128, 64, 1, 1, 1, 128, 512, 1024, 1, 7168, 512, 1024, 9, 4096, 1, "MN-major M128 N64 halo wgrad, A start +4096 B"},
      {128, 64, 1, 1, 1, 128, 512, 1024, 1, 7168, 512, 1024, 9, 0, 3, "MN-major M128 N64 halo wgrad, 3 accumulators in turn"},
      {128, 64, 1, 1, 1, 128, 512, 1024, 1, 7168, 512, 1024, 9, 3584, 3, "MN-major M128 N64 halo wgrad, +3584 B and 3 accumulators"},
      {128, 64, 1, 1, 1, 9216, 512, 1024, 1, 7168, 512, 1024, 9, 512, 1, "MN-major M128 N64 groups 9 KB apart, A start +512 B"},
      {128, 64, 0, 0, 2, 16, 1024, 32, 2, 16, 1024, 32, 4, 128, 1, "K-major sw128 M128 N64, A start +128 B (halo fprop row shift)"},
      {128, 64, 0, 0, 2, 16, 1024, 32, 2, 16, 1024, 32, 4, 384, 1, "K-major sw128 M128 N64, A start +384 B"},
  };
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int R = 2048;
  for (const Cfg& c : cfgs) {
    long long h[2] = {0, 0};
    double res[8];
    for (int poll : {0, 2, 4, 6}) {
      for (int rep = 0; rep < 2; ++rep) {
        rate_kernel<<<(poll & 4) ? 148 : 1, 192, 200 * 1024>>>(c, R, out, poll & 3);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      }
      res[poll] = (double)h[1] / R;
    }
    printf("%-70s %6.1f cyc/MMA zeros 1 CTA, %6.1f non-zero 1 CTA, %6.1f zeros 148 CTAs, %6.1f non-zero 148 CTAs\n", c.name, res[0], res[2], res[4], res[6]);
  }
  return 0;
}
