// Stand-alone probe of the tcgen05/TMA primitives in scanet3_b200/csrc/tc_sm100.cuh (developer tool, not a test).
// Single CTA: TMA-load A[128 x 32] (K-major) and B (K-major [N x 32] or MN-major [32 x N]), 4 MMAs (K = 8 each), read the
// accumulator back from TMEM and compare with a CPU product.   build: see tests/probe/run.sh
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../scanet3_b200/csrc/tc_sm100.cuh"

using namespace sb::tc;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at line %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap map2d(const float* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer, int atom32 = 0) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  CUtensorMap m;
  cuuint64_t gdim[2] = {inner, outer}, gstr[1] = {inner * 4};
  cuuint32_t box[2] = {box_inner, box_outer}, es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

// mode 0: B K-major [N rows x 32] ; mode 1: B MN-major [32 k-rows x N], loaded as N/32 boxes of [32 x 32]
// a_mn = 1: A MN-major [32 k-rows x 128 m] loaded as 4 boxes
__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                       float* D, float* dump, int N, int mode, int a_mn, int M, int lay_mn, int sbo_mn, int kstep_mn) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;
  uint8_t* sb = smem + 16384;
  uint64_t* bar = (uint64_t*)(smem + 16384 + 32768);
  uint64_t* done = bar + 1;
  uint32_t* slot = (uint32_t*)(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t a_bytes = 16384 * (a_mn ? M / 128 : 1) * (a_mn ? 1 : 1);
    mbar_arrive_expect_tx(bar, (a_mn ? (M / 32) * 4096 : 16384) + N * 128);
    (void)a_bytes;
    if (a_mn) {
      for (int j = 0; j < M / 32; ++j) tma_load_2d(sa + j * 4096, &map_a, bar, j * 32, 0);
    } else {
      tma_load_2d(sa, &map_a, bar, 0, 0);
    }
    if (mode == 1) {
      for (int j = 0; j < N / 32; ++j) tma_load_2d(sb + j * 4096, &map_b, bar, j * 32, 0);
    } else {
      tma_load_2d(sb, &map_b, bar, 0, 0);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = idesc_tf32(M, N, a_mn, mode);
    for (int ks = 0; ks < 4; ++ks) {
      uint64_t da = a_mn ? smem_desc(smem_u32(sa) + ks * kstep_mn, 4096, sbo_mn, lay_mn) : smem_desc_sw128(smem_u32(sa) + ks * 32, 16, 1024);
      uint64_t db = mode == 1 ? smem_desc(smem_u32(sb) + ks * kstep_mn, 4096, sbo_mn, lay_mn) : smem_desc_sw128(smem_u32(sb) + ks * 32, 16, 1024);
      umma_tf32(tmem, da, db, idesc, ks ? 1u : 0u);
    }
    umma_commit(done);
  }
  __syncthreads();
  // raw smem dump of the A tile (first 2 KB) to see what TMA wrote
  for (int i = threadIdx.x; i < 512; i += blockDim.x) dump[i] = ((float*)sa)[i];
  mbar_wait(done, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0;
  int N = argc > 2 ? atoi(argv[2]) : 64;
  int a_mn = argc > 3 ? atoi(argv[3]) : 0;
  int M = argc > 4 ? atoi(argv[4]) : 128;
  int atom32 = argc > 5 ? atoi(argv[5]) : 0;
  int sbo_mn = argc > 6 ? atoi(argv[6]) : 1024;
  int kstep_mn = argc > 7 ? atoi(argv[7]) : 1024;
  const int K = 32;
  std::vector<float> A(M * K), B(N * K), Dh(128 * N, -777.f), ref(M * N);
  for (int i = 0; i < M * K; ++i) A[i] = (float)((i * 7 + 3) % 13 - 6);  // logical A[m][k], m = i / K
  for (int i = 0; i < N * K; ++i) B[i] = (float)((i * 5 + 1) % 11 - 5);  // logical B[n][k], n = i / K
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
      ref[m * N + n] = s;
    }
  std::vector<float> Bmem(N * K), Amem(M * K);
  if (mode == 1) for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bmem[k * N + n] = B[n * K + k];  // [K rows][N]
  else Bmem = B;
  if (a_mn) for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) Amem[k * M + m] = A[m * K + k];  // [K rows][M]
  else Amem = A;
  float *dA, *dB, *dD, *dDump;
  CK(cudaMalloc(&dA, Amem.size() * 4)); CK(cudaMalloc(&dB, Bmem.size() * 4)); CK(cudaMalloc(&dD, Dh.size() * 4)); CK(cudaMalloc(&dDump, 2048));
  CK(cudaMemcpy(dA, Amem.data(), Amem.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bmem.data(), Bmem.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dD, Dh.data(), Dh.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap ma = a_mn ? map2d(dA, M, K, 32, 32, atom32) : map2d(dA, K, M, 32, 128);
  CUtensorMap mb = mode == 1 ? map2d(dB, N, K, 32, 32, atom32) : map2d(dB, K, N, 32, N);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000));
  probe_kernel<<<1, 128, 60000>>>(ma, mb, dD, dDump, N, mode, a_mn, M, atom32 ? 1 : 2, sbo_mn, kstep_mn);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<float> dump(512);
  CK(cudaMemcpy(dump.data(), dDump, 2048, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int bad = 0;
  for (int m = 0; m < M; ++m) {
    int lane_row = (M == 64) ? ((m % 16) + 32 * (m / 16)) : m;
    for (int n = 0; n < N; ++n) {
      double e = fabs(Dh[lane_row * N + n] - ref[m * N + n]);
      if (e > maxerr) maxerr = e;
      if (e > 1e-3) ++bad;
    }
  }
  printf("atom32=%d sbo=%d kstep=%d ", atom32, sbo_mn, kstep_mn);
  printf("mode=%d N=%d a_mn=%d M=%d: max abs err %.4f, bad %d / %d\n", mode, N, a_mn, M, maxerr, bad, M * N);
  printf("D[0][0..7]   = "); for (int j = 0; j < 8; ++j) printf("%8.1f", Dh[j]); printf("\n");
  printf("ref[0][0..7] = "); for (int j = 0; j < 8; ++j) printf("%8.1f", ref[j]); printf("\n");
  printf("D[1][0..7]   = "); for (int j = 0; j < 8; ++j) printf("%8.1f", Dh[N + j]); printf("\n");
  printf("ref[1][0..7] = "); for (int j = 0; j < 8; ++j) printf("%8.1f", ref[N + j]); printf("\n");
  printf("smem A row0 = "); for (int j = 0; j < 32; ++j) printf("%4.0f", dump[j]); printf("\n");
  printf("smem A row1 = "); for (int j = 0; j < 32; ++j) printf("%4.0f", dump[32 + j]); printf("\n");
  printf("A[0][0..31] = "); for (int j = 0; j < 32; ++j) printf("%4.0f", A[j]); printf("\n");
  return bad ? 2 : 0;
}
