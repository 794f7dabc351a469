// Vectorised HBM/L2-bound kernels for the shapes the BASELINE configs hit (128-bit accesses, channel-fastest NHWC):
//   * max-pool forward / backward (gather form, first-max rule) with the neighbouring ReLU folded into the backward,
//   * Conv2D with a tiny reduction (KH*KW*Cin <= 32, e.g. the first layer on 1- or 3-channel images): fprop (+ReLU) and wgrad,
//   * the skinny dense head (N <= 16 outputs): split-K forward, fused bias + softmax + CCE (+ bias gradient), wgrad, dgrad.
// All reductions use fixed orders (no atomics), so results are run-to-run deterministic.
#include "common.cuh"
#include "kernels.cuh"

namespace sb {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 max4(float4 a, float4 b) { return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)); }

static int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  long long cap = (long long)kNumSMs * 32;
  return (int)(b > cap ? cap : b);
}

// =====================================================================================================================
// max pool, C % 4 == 0  (math/nn/AllKernels.scala:270-324)
// =====================================================================================================================
__global__ void __launch_bounds__(256) pool_max_fwd_v4_kernel(const float* __restrict__ x, float* __restrict__ y, PoolGeom g, long long n4) {
  const int C4 = g.C >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long long t = i / C4;
    const int ow = (int)(t % g.OW); t /= g.OW;
    const int oh = (int)(t % g.OH);
    const int b = (int)(t / g.OH);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int a = 0; a < g.KH; ++a) {
      const int ih = oh * g.SH + a - g.PT;
      if (ih < 0 || ih >= g.H) continue;
      for (int bb = 0; bb < g.KW; ++bb) {
        const int iw = ow * g.SW + bb - g.PL;
        if (iw < 0 || iw >= g.W) continue;
        m = max4(m, ld4(x + (((long long)b * g.H + ih) * g.W + iw) * g.C + c4 * 4));
      }
    }
    st4(y + i * 4, m);
  }
}
void pool_max_fwd_v4(const float* x, float* y, const PoolGeom& g, cudaStream_t s) {
  const long long n4 = (long long)g.N * g.OH * g.OW * (g.C / 4);
  pool_max_fwd_v4_kernel<<<grid_for(n4, 256), 256, 0, s>>>(x, y, g, n4);
  SB_LAUNCHED();
}

// backward, gather form: input element (h,w) receives dy of every window whose FIRST maximum (row-major) it is.
// RELU: x is the output of an in-place ReLU(thr) and the result is additionally masked with [x > thr].
__device__ __forceinline__ void pool_win_acc(float self, float v, bool earlier, bool& win) {
  if (earlier ? (v >= self) : (v > self)) win = false;
}
template <bool RELU>
__global__ void __launch_bounds__(256) pool_max_bwd_v4_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              float* __restrict__ dx, PoolGeom g, float thr, long long n4) {
  const int C4 = g.C >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long long t = i / C4;
    const int w = (int)(t % g.W); t /= g.W;
    const int h = (int)(t % g.H);
    const int b = (int)(t / g.H);
    const float* xb = x + (long long)b * g.H * g.W * g.C + c4 * 4;
    const float4 self = ld4(xb + ((long long)h * g.W + w) * g.C);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int oh_lo = h + g.PT - g.KH + 1; oh_lo = oh_lo <= 0 ? 0 : (oh_lo + g.SH - 1) / g.SH;
    int oh_hi = (h + g.PT) / g.SH; if (oh_hi > g.OH - 1) oh_hi = g.OH - 1;
    int ow_lo = w + g.PL - g.KW + 1; ow_lo = ow_lo <= 0 ? 0 : (ow_lo + g.SW - 1) / g.SW;
    int ow_hi = (w + g.PL) / g.SW; if (ow_hi > g.OW - 1) ow_hi = g.OW - 1;
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        bool w0 = true, w1 = true, w2 = true, w3 = true;
        for (int a = 0; a < g.KH; ++a) {
          const int ih = oh * g.SH + a - g.PT;
          if (ih < 0 || ih >= g.H) continue;
          for (int bb = 0; bb < g.KW; ++bb) {
            const int iw = ow * g.SW + bb - g.PL;
            if (iw < 0 || iw >= g.W || (ih == h && iw == w)) continue;
            const float4 v = ld4(xb + ((long long)ih * g.W + iw) * g.C);
            const bool earlier = ih < h || (ih == h && iw < w);
            pool_win_acc(self.x, v.x, earlier, w0);
            pool_win_acc(self.y, v.y, earlier, w1);
            pool_win_acc(self.z, v.z, earlier, w2);
            pool_win_acc(self.w, v.w, earlier, w3);
          }
        }
        const float4 gy = ld4(dy + (((long long)b * g.OH + oh) * g.OW + ow) * g.C + c4 * 4);
        if (w0) acc.x = __fadd_rn(acc.x, gy.x);
        if (w1) acc.y = __fadd_rn(acc.y, gy.y);
        if (w2) acc.z = __fadd_rn(acc.z, gy.z);
        if (w3) acc.w = __fadd_rn(acc.w, gy.w);
      }
    if (RELU) {
      if (!(self.x > thr)) acc.x = 0.f;
      if (!(self.y > thr)) acc.y = 0.f;
      if (!(self.z > thr)) acc.z = 0.f;
      if (!(self.w > thr)) acc.w = 0.f;
    }
    st4(dx + i * 4, acc);
  }
}
void pool_max_bwd_v4(const float* x, const float* dy, float* dx, const PoolGeom& g, int relu, float thr, cudaStream_t s) {
  const long long n4 = (long long)g.N * g.H * g.W * (g.C / 4);
  if (relu) pool_max_bwd_v4_kernel<true><<<grid_for(n4, 256), 256, 0, s>>>(x, dy, dx, g, thr, n4);
  else pool_max_bwd_v4_kernel<false><<<grid_for(n4, 256), 256, 0, s>>>(x, dy, dx, g, 0.f, n4);
  SB_LAUNCHED();
}

// =====================================================================================================================
// Conv2D with a tiny reduction: K = KH*KW*Cin <= 32, Cout % 4 == 0 (any stride / padding)
// =====================================================================================================================
constexpr int kSmallK = 32;
__global__ void __launch_bounds__(256) conv_smallk_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              float* __restrict__ y, ConvGeom g, int relu, float thr, long long n4) {
  const int C4 = g.Cout >> 2;
  const int K = g.KH * g.KW * g.Cin;
  __shared__ float sw[kSmallK * 256];  // filter [K][Cout], Cout <= 256
  for (int i = threadIdx.x; i < K * g.Cout; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long long t = i / C4;
    const int ow = (int)(t % g.OW); t /= g.OW;
    const int oh = (int)(t % g.OH);
    const int b = (int)(t / g.OH);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = 0;
    for (int a = 0; a < g.KH; ++a) {
      const int ih = oh * g.SH + a - g.PT;
      for (int bb = 0; bb < g.KW; ++bb) {
        const int iw = ow * g.SW + bb - g.PL;
        const bool ok = ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
        const float* xp = x + (((long long)b * g.H + ih) * g.W + iw) * g.Cin;
        for (int ci = 0; ci < g.Cin; ++ci, ++k) {
          const float xv = ok ? __ldg(xp + ci) : 0.f;
          const float4 wv = *reinterpret_cast<const float4*>(&sw[k * g.Cout + c4 * 4]);
          acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y);
          acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
        }
      }
    }
    if (relu) { acc.x = fmaxf(thr, acc.x); acc.y = fmaxf(thr, acc.y); acc.z = fmaxf(thr, acc.z); acc.w = fmaxf(thr, acc.w); }
    st4(y + i * 4, acc);
  }
}
void conv_smallk_fwd(const float* x, const float* w, float* y, const ConvGeom& g, int relu, float thr, cudaStream_t s) {
  const long long n4 = (long long)g.N * g.OH * g.OW * (g.Cout / 4);
  conv_smallk_fwd_kernel<<<grid_for(n4, 256), 256, 0, s>>>(x, w, y, g, relu, thr, n4);
  SB_LAUNCHED();
}

// wgrad: dW[k][co] = sum_pixels X(pixel, k) * dY[pixel][co].  Block = (Cout/4) channel groups x PL pixel lanes; every thread
// keeps K float4 accumulators; lanes are tree-free summed in lane order through shared memory; one partial per block.
template <int KH, int KW, int CIN>
__global__ void __launch_bounds__(256) conv_smallk_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                float* __restrict__ partial, ConvGeom g, long long pixels,
                                                                long long pix_per_block) {
  constexpr int K = KH * KW * CIN;
  const int C4 = g.Cout >> 2;
  const int lanes = blockDim.x / C4;
  const int c4 = threadIdx.x % C4, lane = threadIdx.x / C4;
  float4 acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  const long long p1 = p0 + pix_per_block < pixels ? p0 + pix_per_block : pixels;
  for (long long p = p0 + lane; p < p1; p += lanes) {
    const int ow = (int)(p % g.OW);
    long long t = p / g.OW;
    const int oh = (int)(t % g.OH);
    const int b = (int)(t / g.OH);
    const float4 gy = ld4(dy + p * g.Cout + c4 * 4);
#pragma unroll
    for (int a = 0; a < KH; ++a) {
      const int ih = oh * g.SH + a - g.PT;
#pragma unroll
      for (int bb = 0; bb < KW; ++bb) {
        const int iw = ow * g.SW + bb - g.PL;
        const bool ok = ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
        const float* xp = x + (((long long)b * g.H + ih) * g.W + iw) * CIN;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float xv = ok ? __ldg(xp + ci) : 0.f;
          float4& r = acc[(a * KW + bb) * CIN + ci];
          r.x = fmaf(xv, gy.x, r.x); r.y = fmaf(xv, gy.y, r.y); r.z = fmaf(xv, gy.z, r.z); r.w = fmaf(xv, gy.w, r.w);
        }
      }
    }
  }
  // cross-lane sum in lane order, one k at a time through shared memory
  __shared__ float4 red[256];
  float* part = partial + (long long)blockIdx.x * K * g.Cout;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    red[threadIdx.x] = acc[k];
    __syncthreads();
    if (lane == 0) {
      float4 tot = red[c4];
      for (int l = 1; l < lanes; ++l) {
        const float4 v = red[l * C4 + c4];
        tot.x = __fadd_rn(tot.x, v.x); tot.y = __fadd_rn(tot.y, v.y); tot.z = __fadd_rn(tot.z, v.z); tot.w = __fadd_rn(tot.w, v.w);
      }
      st4(part + k * g.Cout + c4 * 4, tot);
    }
    __syncthreads();
  }
}
__global__ void partial_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int n, int parts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = partial[i];
  for (int c = 1; c < parts; ++c) acc = __fadd_rn(acc, partial[(size_t)c * n + i]);
  out[i] = acc;
}
// filter shapes 3x3x1, 3x3x3, 5x5x1 (fully unrolled accumulators); ws must hold blocks*K*Cout floats
bool conv_smallk_wgrad(const float* x, const float* dy, float* dw, const ConvGeom& g, float* ws, size_t ws_floats, cudaStream_t s) {
  const int K = g.KH * g.KW * g.Cin;
  const int C4 = g.Cout / 4;
  const int shape = g.KH * 100 + g.KW * 10 + g.Cin;
  if ((shape != 331 && shape != 333 && shape != 551) || C4 > 256 || 256 % C4) return false;
  const long long pixels = (long long)g.N * g.OH * g.OW;
  int blocks = 2 * kNumSMs;
  if ((size_t)blocks * K * g.Cout > ws_floats) blocks = (int)(ws_floats / ((size_t)K * g.Cout));
  if (blocks < 1) return false;
  if (blocks > pixels) blocks = (int)pixels;
  const long long ppb = (pixels + blocks - 1) / blocks;
  blocks = (int)((pixels + ppb - 1) / ppb);
  if (shape == 331) conv_smallk_wgrad_kernel<3, 3, 1><<<blocks, 256, 0, s>>>(x, dy, ws, g, pixels, ppb);
  else if (shape == 333) conv_smallk_wgrad_kernel<3, 3, 3><<<blocks, 256, 0, s>>>(x, dy, ws, g, pixels, ppb);
  else conv_smallk_wgrad_kernel<5, 5, 1><<<blocks, 256, 0, s>>>(x, dy, ws, g, pixels, ppb);
  SB_LAUNCHED();
  const int n = K * g.Cout;
  partial_reduce_kernel<<<(n + 255) / 256, 256, 0, s>>>(ws, dw, n, blocks);
  SB_LAUNCHED();
  return true;
}

// =====================================================================================================================
// skinny dense head: Z[M,N] = X[M,K] . W[K,N], N <= 16 (layer/Dense.scala:57-61 feeding Bias + Softmax + CCE)
// =====================================================================================================================
constexpr int kSkinnyN = 16;
constexpr int kSkinnyKC = 256;  // K rows per block
// partial[chunk][m][n]: block = one K chunk; warps stride over rows m; lanes stride over k inside the chunk
__global__ void __launch_bounds__(256) dense_skinny_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               float* __restrict__ partial, int M, int K, int N) {
  __shared__ float sw[kSkinnyKC * kSkinnyN];
  const int k0 = blockIdx.x * kSkinnyKC;
  const int kc = min(kSkinnyKC, K - k0);
  for (int i = threadIdx.x; i < kc * N; i += blockDim.x) sw[i] = w[(size_t)k0 * N + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int m = warp; m < M; m += nw) {
    float acc[kSkinnyN];
#pragma unroll
    for (int n = 0; n < kSkinnyN; ++n) acc[n] = 0.f;
    const float* xr = x + (size_t)m * K + k0;
    for (int k = lane; k < kc; k += 32) {
      const float xv = __ldg(xr + k);
#pragma unroll
      for (int n = 0; n < kSkinnyN; ++n)
        if (n < N) acc[n] = fmaf(xv, sw[k * N + n], acc[n]);
    }
#pragma unroll
    for (int n = 0; n < kSkinnyN; ++n) {
      if (n < N) {
        float v = acc[n];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) partial[((size_t)blockIdx.x * M + m) * N + n] = v;
      }
    }
  }
}
__global__ void skinny_reduce_kernel(const float* __restrict__ partial, float* __restrict__ z, int mn, int chunks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mn) return;
  float acc = partial[i];
  for (int c = 1; c < chunks; ++c) acc = __fadd_rn(acc, partial[(size_t)c * mn + i]);
  z[i] = acc;
}
bool dense_skinny_fwd(const float* x, const float* w, float* z, int M, int N, int K, float* ws, size_t ws_floats, cudaStream_t s) {
  if (N > kSkinnyN) return false;
  const int chunks = (K + kSkinnyKC - 1) / kSkinnyKC;
  if ((size_t)chunks * M * N > ws_floats) return false;
  dense_skinny_fwd_kernel<<<chunks, 256, 0, s>>>(x, w, ws, M, K, N);
  SB_LAUNCHED();
  skinny_reduce_kernel<<<(M * N + 255) / 256, 256, 0, s>>>(ws, z, M * N, chunks);
  SB_LAUNCHED();
  return true;
}
// dW[k][n] = sum_m X[m][k] * G[m][n]: one thread per k, G staged in shared memory
__global__ void __launch_bounds__(256) dense_skinny_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gmat,
                                                                 float* __restrict__ dw, int M, int K, int N) {
  extern __shared__ float sg[];  // [M][N]
  for (int i = threadIdx.x; i < M * N; i += blockDim.x) sg[i] = gmat[i];
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float acc[kSkinnyN];
#pragma unroll
  for (int n = 0; n < kSkinnyN; ++n) acc[n] = 0.f;
  for (int m = 0; m < M; ++m) {
    const float xv = __ldg(x + (size_t)m * K + k);
#pragma unroll
    for (int n = 0; n < kSkinnyN; ++n)
      if (n < N) acc[n] = fmaf(xv, sg[m * N + n], acc[n]);
  }
#pragma unroll
  for (int n = 0; n < kSkinnyN; ++n)
    if (n < N) dw[(size_t)k * N + n] = acc[n];
}
bool dense_skinny_wgrad(const float* x, const float* g, float* dw, int M, int N, int K, cudaStream_t s) {
  if (N > kSkinnyN || (size_t)M * N * 4 > 40000) return false;
  dense_skinny_wgrad_kernel<<<(K + 255) / 256, 256, (size_t)M * N * 4, s>>>(x, g, dw, M, K, N);
  SB_LAUNCHED();
  return true;
}
// dX[m][k] = sum_n G[m][n] * W[k][n]
__global__ void __launch_bounds__(256) dense_skinny_dgrad_kernel(const float* __restrict__ gmat, const float* __restrict__ w,
                                                                 float* __restrict__ dx, int M, int K, int N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * K) return;
  const int m = (int)(i / K), k = (int)(i % K);
  float acc = 0.f;
  for (int n = 0; n < N; ++n) acc = fmaf(__ldg(gmat + m * N + n), __ldg(w + (size_t)k * N + n), acc);
  dx[i] = acc;
}
bool dense_skinny_dgrad(const float* g, const float* w, float* dx, int M, int N, int K, cudaStream_t s) {
  if (N > kSkinnyN) return false;
  const long long n = (long long)M * K;
  dense_skinny_dgrad_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(g, w, dx, M, K, N);
  SB_LAUNCHED();
  return true;
}

// Fused head: z += bias ; p = exp(z)/sum ; loss = -sum(y log(p+eps))/B ; dz = -(p/B)(r - sum r p), r = y/(p+eps) ;
// db = column sums of dz (fixed row order).  One CTA; rows over warps.  (layer/Bias.scala:32-33, models/Activation.scala:63-69,
// models/Loss.scala:44-57)
__global__ void __launch_bounds__(1024) head_bias_softmax_cce_kernel(float* __restrict__ z, const float* __restrict__ bias,
                                                                     const float* __restrict__ y, float* __restrict__ dz,
                                                                     float* __restrict__ dbias, StepState* st, int rows, int C) {
  const float eps = 1e-8f, fr = (float)rows;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __shared__ float sl[32];
  float acc = 0.f;
  for (int row = wid; row < rows; row += nw) {
    float* zr = z + (size_t)row * C;
    const float* yr = y + (size_t)row * C;
    float e = 0.f, yv = 0.f;
    if (lane < C) {
      e = expf(__fadd_rn(zr[lane], bias ? bias[lane] : 0.f));
      yv = yr[lane];
    }
    float sum = e;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
    const float p = __fdiv_rn(e, sum);
    const float pe = __fadd_rn(p, eps);
    float l = lane < C ? __fmul_rn(yv, logf(pe)) : 0.f;
    const float r = lane < C ? __fdiv_rn(yv, pe) : 0.f;
    float dot = __fmul_rn(r, p);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l = __fadd_rn(l, __shfl_xor_sync(0xffffffffu, l, o));
      dot = __fadd_rn(dot, __shfl_xor_sync(0xffffffffu, dot, o));
    }
    acc = __fadd_rn(acc, l);
    if (lane < C) {
      zr[lane] = p;
      if (dz) dz[(size_t)row * C + lane] = __fmul_rn(-__fdiv_rn(p, fr), __fsub_rn(r, dot));
    }
  }
  if (lane == 0) sl[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < nw; ++i) tot = __fadd_rn(tot, sl[i]);
    st->loss = __fdiv_rn(-tot, fr);
  }
  __syncthreads();  // dz rows of every warp are visible
  if (dz && dbias && threadIdx.x < C) {
    float s = 0.f;
    for (int row = 0; row < rows; ++row) s = __fadd_rn(s, dz[(size_t)row * C + threadIdx.x]);
    dbias[threadIdx.x] = s;
  }
}
bool head_bias_softmax_cce(float* z, const float* bias, const float* y, float* dz, float* dbias, StepState* st, int rows, int C,
                           cudaStream_t s) {
  if (C > 32) return false;
  head_bias_softmax_cce_kernel<<<1, 1024, 0, s>>>(z, bias, y, dz, dbias, st, rows, C);
  SB_LAUNCHED();
  return true;
}

}  // namespace sb
