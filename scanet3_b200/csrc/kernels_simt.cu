// fp32 FFMA kernels of the scanet-b200 engine (sm_100a): the parity-mode contractions and every
// bandwidth-bound kernel of the train step.  Math follows the reference's op graph; the cited lines are
// under /root/reference/src/main/scala/scanet/.  Element-wise formulas use __fmul_rn/__fadd_rn where the
// reference executes separate TF ops, so that nvcc does not contract them into FMAs.
#include "common.cuh"
#include "kernels.cuh"

namespace sb {

thread_local long long tl_launch_count = 0;

// =================================================================================================
// step prologue
// =================================================================================================
template <typename V>
__global__ void step_prologue_kernel(StepState* st, const V* __restrict__ ds_x, const V* __restrict__ ds_y,
                                     V* __restrict__ bx, V* __restrict__ by, long long x4, long long y4,
                                     float beta1, float beta2, int batch_from_state, float* __restrict__ prev_loss_slot, int early) {
  pdl_prologue_trigger_if(early);
  // resident shard: batch index from the device state (only the optimizer kernel of the previous step writes it);
  // host-fed staging buffer: always its single batch
  const long long b = batch_from_state ? st->batch_idx : 0;
  if (ds_x != nullptr) {
    const V* sx = ds_x + b * x4;
    const V* sy = ds_y + b * y4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < x4 + y4; i += stride) {
      if (i < x4) bx[i] = sx[i];
      else by[i - x4] = sy[i - x4];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // the loss of the step BEFORE this one (steps chained on the device): a fixed slot, or -- batches taken by the device-side
    // cursor -- the slot of the previous batch
    if (prev_loss_slot) {
      if (!batch_from_state) *prev_loss_slot = st->loss;
      else if (b > 0) prev_loss_slot[b - 1] = st->loss;
    }
    // t = globalIter + localIter + 1, fed as int32 and cast to float (optimizers/Optimizer.scala:143);
    // boost(x, rate, t) = x / (1 - rate^float(t)) (math/alg/AllKernels.scala:546-549).  The power is
    // evaluated in double and rounded once, i.e. the correctly rounded powf TF-CPU (glibc) returns.
    const long long t = st->iter + 1;
    if (st->bc_iter == t) {  // computed one step ahead by the optimizer kernel (same expression, same bits)
      st->bc1 = st->bc1_next;
      st->bc2 = st->bc2_next;
    } else {
      const double tf = (double)(float)t;
      st->bc1 = __fsub_rn(1.0f, (float)pow((double)beta1, tf));
      st->bc2 = __fsub_rn(1.0f, (float)pow((double)beta2, tf));
    }
    st->iter = t;
  }
}

void step_prologue(StepState* st, const float* ds_x, const float* ds_y, float* bx, float* by, long long x_per_batch,
                   long long y_per_batch, float beta1, float beta2, int batch_from_state, cudaStream_t s, float* prev_loss_slot) {
  // 128-bit copies when every batch of the shard starts 16-byte aligned, scalar copies otherwise
  const bool vec = (x_per_batch % 4 == 0) && (y_per_batch % 4 == 0);
  const long long xn = vec ? x_per_batch / 4 : x_per_batch, yn = vec ? y_per_batch / 4 : y_per_batch;
  int blocks = 1;
  if (ds_x) blocks = (int)fmin((double)ceil_div(xn + yn, 256), (double)kNumSMs * 8);
  if (blocks < 1) blocks = 1;
  if (vec)
    launch_k(step_prologue_kernel<float4>, blocks, 256, 0, s, st, (const float4*)ds_x, (const float4*)ds_y, (float4*)bx,
                                                        (float4*)by, xn, yn, beta1, beta2, batch_from_state, prev_loss_slot, pdl_early_hint());
  else
    launch_k(step_prologue_kernel<float>, blocks, 256, 0, s, st, ds_x, ds_y, bx, by, xn, yn, beta1, beta2, batch_from_state,
             prev_loss_slot, pdl_early_hint());
  SB_LAUNCHED();
}

__global__ void store_loss_kernel(const StepState* __restrict__ st, float* __restrict__ slot, int by_cursor) {
  pdl_prologue();
  if (!by_cursor) *slot = st->loss;
  else if (st->batch_idx > 0) slot[st->batch_idx - 1] = st->loss;  // the batch the device-side cursor just left
}
void store_loss(const StepState* st, float* slot, cudaStream_t s, int by_cursor) {
  launch_k(store_loss_kernel, 1, 1, 0, s, st, slot, by_cursor);
  SB_LAUNCHED();
}

// =================================================================================================
// generic tiled fp32 GEMM:  C[M,N] = sum_k A(m,k) * B(k,n), operands addressed through loaders so the same
// kernel serves MatMul NN/NT/TN (math/alg/AllKernels.scala:370-389) and Conv2D fwd / dgrad / wgrad as
// implicit GEMMs (math/nn/AllKernels.scala:143-261).  64x64 tile, BK=16, 256 threads, 4x4 per thread.
// Deterministic split-K: partial tiles go to a workspace and are summed in split order.
// =================================================================================================
struct LdRowMajor {  // X[r*ld + c]
  const float* p;
  int ld;
  __device__ __forceinline__ float operator()(int r, int c) const { return __ldg(p + (long long)r * ld + c); }
};
struct LdColMajor {  // X[c*ld + r]
  const float* p;
  int ld;
  __device__ __forceinline__ float operator()(int r, int c) const { return __ldg(p + (long long)c * ld + r); }
};
// A(m,k) of conv fprop: m=(n,oh,ow), k=(kh,kw,ci)
struct LdConvFwdA {
  const float* x;
  ConvGeom g;
  __device__ __forceinline__ float operator()(int m, int k) const {
    int ow = m % g.OW, t = m / g.OW, oh = t % g.OH, n = t / g.OH;
    int ci = k % g.Cin, t2 = k / g.Cin, kw = t2 % g.KW, kh = t2 / g.KW;
    int ih = oh * g.SH + kh - g.PT, iw = ow * g.SW + kw - g.PL;
    if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) return 0.f;
    return __ldg(x + (((long long)n * g.H + ih) * g.W + iw) * g.Cin + ci);
  }
};
// A(m,k) of conv dgrad: m=(n,h,w), k=(kh,kw,co) -> dY[n,(h+PT-kh)/SH,(w+PL-kw)/SW,co]
struct LdConvDgradA {
  const float* dy;
  ConvGeom g;
  __device__ __forceinline__ float operator()(int m, int k) const {
    int w = m % g.W, t = m / g.W, h = t % g.H, n = t / g.H;
    int co = k % g.Cout, t2 = k / g.Cout, kw = t2 % g.KW, kh = t2 / g.KW;
    int th = h + g.PT - kh, tw = w + g.PL - kw;
    if (th < 0 || tw < 0 || th % g.SH || tw % g.SW) return 0.f;
    int oh = th / g.SH, ow = tw / g.SW;
    if (oh >= g.OH || ow >= g.OW) return 0.f;
    return __ldg(dy + (((long long)n * g.OH + oh) * g.OW + ow) * g.Cout + co);
  }
};
// B(k,n) of conv dgrad: k=(kh,kw,co), n=ci -> W[kh,kw,ci,co]
struct LdConvDgradB {
  const float* w;
  ConvGeom g;
  __device__ __forceinline__ float operator()(int k, int ci) const {
    int co = k % g.Cout, t2 = k / g.Cout;  // t2 = kh*KW+kw
    return __ldg(w + ((long long)t2 * g.Cin + ci) * g.Cout + co);
  }
};
// A(m,k) of conv wgrad: m=(kh,kw,ci), k=(n,oh,ow) -> X[n,oh*SH+kh-PT,ow*SW+kw-PL,ci]
struct LdConvWgradA {
  const float* x;
  ConvGeom g;
  __device__ __forceinline__ float operator()(int m, int k) const {
    int ci = m % g.Cin, t2 = m / g.Cin, kw = t2 % g.KW, kh = t2 / g.KW;
    int ow = k % g.OW, t = k / g.OW, oh = t % g.OH, n = t / g.OH;
    int ih = oh * g.SH + kh - g.PT, iw = ow * g.SW + kw - g.PL;
    if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) return 0.f;
    return __ldg(x + (((long long)n * g.H + ih) * g.W + iw) * g.Cin + ci);
  }
};

template <class AL, class BL, bool A_K_CONTIG, bool B_N_CONTIG>
__global__ void __launch_bounds__(256) gemm_tiled_kernel(AL a, BL b, float* __restrict__ C, int M, int N, int K,
                                                         int k_per_split, float* __restrict__ partial) {
  pdl_prologue();
  __shared__ float As[16][68];
  __shared__ float Bs[16][68];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int kb = blockIdx.z * k_per_split;
  const int ke = min(K, kb + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kb; k0 < ke; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int m, k;
      if (A_K_CONTIG) { m = e >> 4; k = e & 15; } else { k = e >> 6; m = e & 63; }
      int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < ke) ? a(gm, gk) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + i * 256;
      int n, k;
      if (B_N_CONTIG) { k = e >> 6; n = e & 63; } else { n = e >> 4; k = e & 15; }
      int gn = n0 + n, gk = k0 + k;
      Bs[k][n] = (gn < N && gk < ke) ? b(gk, gn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partial ? partial + (long long)blockIdx.z * M * N : C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn < N) out[(long long)gm * N + gn] = acc[i][j];
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ C, long long mn, int splits) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mn) return;
  float acc = partial[i];
  for (int z = 1; z < splits; ++z) acc = __fadd_rn(acc, partial[(long long)z * mn + i]);
  C[i] = acc;
}

template <class AL, class BL, bool A_K_CONTIG, bool B_N_CONTIG>
static void launch_gemm(AL a, BL b, float* C, int M, int N, int K, float* ws, size_t ws_floats, cudaStream_t s) {
  if (M <= 0 || N <= 0) return;
  const int mt = ceil_div(M, 64), nt = ceil_div(N, 64);
  int splits = 1;
  const long long tiles = (long long)mt * nt;
  if (tiles < 2 * kNumSMs && K >= 256) {
    splits = (int)((2 * kNumSMs + tiles - 1) / tiles);
    int max_by_k = K / 64;
    if (splits > max_by_k) splits = max_by_k;
    long long max_by_ws = ws ? (long long)(ws_floats / ((size_t)M * N)) : 1;
    if (splits > max_by_ws) splits = (int)max_by_ws;
    if (splits < 1) splits = 1;
  }
  int kps = ceil_div(K, splits);
  kps = ((kps + 15) / 16) * 16;
  splits = ceil_div(K, kps);
  dim3 grid(mt, nt, splits);
  launch_k(gemm_tiled_kernel<AL, BL, A_K_CONTIG, B_N_CONTIG>, grid, 256, 0, s, a, b, C, M, N, K, kps, splits > 1 ? ws : nullptr);
  SB_LAUNCHED();
  if (splits > 1) {
    long long mn = (long long)M * N;
    launch_k(splitk_reduce_kernel, ceil_div(mn, 256), 256, 0, s, ws, C, mn, splits);
    SB_LAUNCHED();
  }
}

void gemm_nn(const float* A, const float* B, float* C, int M, int N, int K, float* ws, size_t wsf, cudaStream_t s) {
  launch_gemm<LdRowMajor, LdRowMajor, true, true>(LdRowMajor{A, K}, LdRowMajor{B, N}, C, M, N, K, ws, wsf, s);
}
void gemm_nt(const float* A, const float* B, float* C, int M, int N, int K, float* ws, size_t wsf, cudaStream_t s) {
  // B given as [N,K] row-major: B(k,n) = B[n*K+k]
  launch_gemm<LdRowMajor, LdColMajor, true, false>(LdRowMajor{A, K}, LdColMajor{B, K}, C, M, N, K, ws, wsf, s);
}
void gemm_tn(const float* A, const float* B, float* C, int M, int N, int K, float* ws, size_t wsf, cudaStream_t s) {
  // A given as [K,M] row-major: A(m,k) = A[k*M+m]
  launch_gemm<LdColMajor, LdRowMajor, false, true>(LdColMajor{A, M}, LdRowMajor{B, N}, C, M, N, K, ws, wsf, s);
}
void conv2d_fwd(const float* x, const float* w, float* y, const ConvGeom& g, float* ws, size_t wsf, cudaStream_t s) {
  launch_gemm<LdConvFwdA, LdRowMajor, true, true>(LdConvFwdA{x, g}, LdRowMajor{w, g.Cout}, y, g.N * g.OH * g.OW, g.Cout,
                                                  g.KH * g.KW * g.Cin, ws, wsf, s);
}
void conv2d_dgrad(const float* dy, const float* w, float* dx, const ConvGeom& g, float* ws, size_t wsf, cudaStream_t s) {
  launch_gemm<LdConvDgradA, LdConvDgradB, true, false>(LdConvDgradA{dy, g}, LdConvDgradB{w, g}, dx, g.N * g.H * g.W, g.Cin,
                                                       g.KH * g.KW * g.Cout, ws, wsf, s);
}
void conv2d_wgrad(const float* x, const float* dy, float* dw, const ConvGeom& g, float* ws, size_t wsf, cudaStream_t s) {
  launch_gemm<LdConvWgradA, LdRowMajor, false, true>(LdConvWgradA{x, g}, LdRowMajor{dy, g.Cout}, dw, g.KH * g.KW * g.Cin,
                                                     g.Cout, g.N * g.OH * g.OW, ws, wsf, s);
}

// =================================================================================================
// bias add / column sums (layer/Bias.scala:32-33; gradient = Sum over all-but-last axes, alg:12-33)
// =================================================================================================
__global__ void bias_add_kernel(float* __restrict__ x, const float* __restrict__ b, long long n, int C) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = __fadd_rn(x[i], __ldg(b + (int)(i % C)));
}
void bias_add(float* x, const float* b, long long rows, int C, cudaStream_t s) {
  long long n = rows * C;
  int blocks = (int)fmin((double)ceil_div(n, 256), (double)kNumSMs * 16);
  launch_k(bias_add_kernel, blocks, 256, 0, s, x, b, n, C);
  SB_LAUNCHED();
}

// column reduction skeleton: block (32 cols, 8 row lanes), grid (col chunks, row chunks);
// F(row, col) -> float2 of two addends; partial[chunk][which][C]
template <class F>
__global__ void __launch_bounds__(256) col_reduce_kernel(F f, float* __restrict__ partial, long long rows, int C,
                                                         long long rows_per_chunk) {
  pdl_prologue();
  __shared__ float sm0[8][33], sm1[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  float a0 = 0.f, a1 = 0.f;
  if (c < C)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      float2 v = f(r, c);
      a0 = __fadd_rn(a0, v.x);
      a1 = __fadd_rn(a1, v.y);
    }
  sm0[threadIdx.y][threadIdx.x] = a0;
  sm1[threadIdx.y][threadIdx.x] = a1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int j = 1; j < 8; ++j) {
      a0 = __fadd_rn(a0, sm0[j][threadIdx.x]);
      a1 = __fadd_rn(a1, sm1[j][threadIdx.x]);
    }
    partial[((long long)blockIdx.y * 2 + 0) * C + c] = a0;
    partial[((long long)blockIdx.y * 2 + 1) * C + c] = a1;
  }
}
struct ColChunks {
  int chunks;
  long long rows_per_chunk;
};
static ColChunks col_chunks(long long rows, int C, size_t ws_floats) {
  int col_blocks = ceil_div(C, 32);
  int want = ceil_div(2 * kNumSMs, col_blocks);
  long long max_by_rows = (rows + 63) / 64;
  if (want > max_by_rows) want = (int)max_by_rows;
  long long max_by_ws = (long long)(ws_floats / ((size_t)2 * C));
  if (want > max_by_ws) want = (int)max_by_ws;
  if (want < 1) want = 1;
  long long rpc = (rows + want - 1) / want;
  return ColChunks{ceil_div(rows, rpc), rpc};
}
template <class F>
static ColChunks launch_col_reduce(F f, float* ws, size_t wsf, long long rows, int C, cudaStream_t s) {
  ColChunks cc = col_chunks(rows, C, wsf);
  dim3 grid(ceil_div(C, 32), cc.chunks);
  launch_k(col_reduce_kernel<F>, grid, dim3(32, 8), 0, s, f, ws, rows, C, cc.rows_per_chunk);
  SB_LAUNCHED();
  return cc;
}
__device__ __forceinline__ float2 finalize_partials(const float* __restrict__ partial, int chunks, int C, int c) {
  float a0 = 0.f, a1 = 0.f;
  for (int j = 0; j < chunks; ++j) {
    a0 = __fadd_rn(a0, partial[((long long)j * 2 + 0) * C + c]);
    a1 = __fadd_rn(a1, partial[((long long)j * 2 + 1) * C + c]);
  }
  return make_float2(a0, a1);
}

struct FSum {
  const float* g;
  int C;
  __device__ __forceinline__ float2 operator()(long long r, int c) const { return make_float2(__ldg(g + r * C + c), 0.f); }
};
__global__ void col_sum_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int chunks, int C) {
  pdl_prologue();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = finalize_partials(partial, chunks, C, c).x;
}
void col_sum(const float* g, float* out, long long rows, int C, float* ws, size_t wsf, cudaStream_t s) {
  ColChunks cc = launch_col_reduce(FSum{g, C}, ws, wsf, rows, C, s);
  launch_k(col_sum_finalize_kernel, ceil_div(C, 128), 128, 0, s, ws, out, cc.chunks, C);
  SB_LAUNCHED();
}

// =================================================================================================
// activations (models/Activation.scala:46-83; local grads math/alg/AllKernels.scala:288-304,391-437)
// act ids: 0 identity, 1 sigmoid, 2 tanh, 3 softmax (row kernels below), 4 relu(thr)
// =================================================================================================
__device__ __forceinline__ float act_apply(float x, int act, float thr) {
  switch (act) {
    case 1: return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
    case 2: return tanhf(x);
    case 4: return fmaxf(thr, x);
    default: return x;
  }
}
// g_in = f'(.) * g_out from the OUTPUT y only: sigmoid (s*(1-s))*g ; tanh (1-y^2)*g ; relu [thr<x]*g == [y>thr]*g
__device__ __forceinline__ float act_grad(float g, float y, int act, float thr) {
  switch (act) {
    case 1: return __fmul_rn(__fmul_rn(y, __fsub_rn(1.f, y)), g);
    case 2: return __fmul_rn(__fsub_rn(1.f, __fmul_rn(y, y)), g);
    case 4: return y > thr ? g : 0.f;
    default: return g;
  }
}
__global__ void act_fwd_kernel(float* __restrict__ x, int act, float thr, long long n) {
  pdl_prologue();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (; i < n; i += stride) {
    if (i + 3 < n) {
      float4 v = *reinterpret_cast<float4*>(x + i);
      v.x = act_apply(v.x, act, thr); v.y = act_apply(v.y, act, thr);
      v.z = act_apply(v.z, act, thr); v.w = act_apply(v.w, act, thr);
      *reinterpret_cast<float4*>(x + i) = v;
    } else {
      for (long long j = i; j < n; ++j) x[j] = act_apply(x[j], act, thr);
    }
  }
}
__global__ void act_bwd_kernel(float* __restrict__ g, const float* __restrict__ y, int act, float thr, long long n) {
  pdl_prologue();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (; i < n; i += stride) {
    if (i + 3 < n) {
      float4 gv = *reinterpret_cast<float4*>(g + i);
      const float4 yv = *reinterpret_cast<const float4*>(y + i);
      gv.x = act_grad(gv.x, yv.x, act, thr); gv.y = act_grad(gv.y, yv.y, act, thr);
      gv.z = act_grad(gv.z, yv.z, act, thr); gv.w = act_grad(gv.w, yv.w, act, thr);
      *reinterpret_cast<float4*>(g + i) = gv;
    } else {
      for (long long j = i; j < n; ++j) g[j] = act_grad(g[j], y[j], act, thr);
    }
  }
}
static int ew_blocks(long long n) {
  long long b = (n / 4 + 255) / 256;
  if (b < 1) b = 1;
  if (b > (long long)kNumSMs * 16) b = (long long)kNumSMs * 16;
  return (int)b;
}
void act_fwd(float* x, int act, float thr, long long n, cudaStream_t s) {
  if (act == 0 || n == 0) return;
  launch_k(act_fwd_kernel, ew_blocks(n), 256, 0, s, x, act, thr, n);
  SB_LAUNCHED();
}
void act_bwd(float* g, const float* y, int act, float thr, long long n, cudaStream_t s) {
  if (act == 0 || n == 0) return;
  launch_k(act_bwd_kernel, ew_blocks(n), 256, 0, s, g, y, act, thr, n);
  SB_LAUNCHED();
}

// softmax rows: e = exp(x) (NO max subtraction), p = e / sum(e)  (models/Activation.scala:63-69)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__global__ void softmax_fwd_kernel(float* __restrict__ x, int rows, int C) {
  pdl_prologue();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* r = x + (long long)row * C;
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) {
    float e = expf(r[c]);
    r[c] = e;
    sum = __fadd_rn(sum, e);
  }
  sum = warp_sum(sum);
  for (int c = lane; c < C; c += 32) r[c] = __fdiv_rn(r[c], sum);
}
// dz = p * (g - sum_j g_j p_j)
__global__ void softmax_bwd_kernel(float* __restrict__ g, const float* __restrict__ p, int rows, int C) {
  pdl_prologue();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* gr = g + (long long)row * C;
  const float* pr = p + (long long)row * C;
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) dot = __fadd_rn(dot, __fmul_rn(gr[c], pr[c]));
  dot = warp_sum(dot);
  for (int c = lane; c < C; c += 32) gr[c] = __fmul_rn(pr[c], __fsub_rn(gr[c], dot));
}
void softmax_fwd(float* x, int rows, int C, cudaStream_t s) {
  launch_k(softmax_fwd_kernel, ceil_div(rows, 8), 256, 0, s, x, rows, C);
  SB_LAUNCHED();
}
void softmax_bwd(float* g, const float* p, int rows, int C, cudaStream_t s) {
  launch_k(softmax_bwd_kernel, ceil_div(rows, 8), 256, 0, s, g, p, rows, C);
  SB_LAUNCHED();
}

// =================================================================================================
// losses (models/Loss.scala:18-57): one CTA, deterministic tree; also writes dLoss/dpred
// =================================================================================================
__device__ float block_sum_1024(float v) {
  __shared__ float sm[32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;  // valid in warp 0
}
__global__ void __launch_bounds__(1024) loss_kernel(int loss, const float* __restrict__ p, const float* __restrict__ y,
                                                    float* __restrict__ dp, StepState* st, long long rows, long long n) {
  pdl_prologue();
  const float eps = 1e-8f;
  const float fn = (float)n, fr = (float)rows;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float pi = p[i], yi = y[i];
    float term, d;
    if (loss == 0) {  // MSE: mean((p-y)^2) ; d = 2(p-y)/n
      float df = __fsub_rn(pi, yi);
      term = __fmul_rn(df, df);
      d = __fdiv_rn(__fmul_rn(2.f, df), fn);
    } else if (loss == 1) {  // BCE: mean(-(y*log(p+eps)) - (1-y)*log(1-(p-eps)))
      float pe = __fadd_rn(pi, eps), q = __fsub_rn(1.f, __fsub_rn(pi, eps)), omy = __fsub_rn(1.f, yi);
      term = __fsub_rn(-__fmul_rn(yi, logf(pe)), __fmul_rn(omy, logf(q)));
      d = __fdiv_rn(__fadd_rn(-__fdiv_rn(yi, pe), __fdiv_rn(omy, q)), fn);
    } else {  // CCE: -sum(y*log(p+eps))/B ; d = -(y/(p+eps))/B
      float pe = __fadd_rn(pi, eps);
      term = __fmul_rn(yi, logf(pe));
      d = __fdiv_rn(-__fdiv_rn(yi, pe), fr);
    }
    acc = __fadd_rn(acc, term);
    if (dp) dp[i] = d;
  }
  float tot = block_sum_1024(acc);
  if (threadIdx.x == 0) st->loss = (loss == 2) ? __fdiv_rn(-tot, fr) : __fdiv_rn(tot, fn);
}
void loss_fwd_bwd(int loss, const float* pred, const float* y, float* dpred, StepState* st, long long rows, long long cols,
                  cudaStream_t s) {
  launch_k(loss_kernel, 1, 1024, 0, s, loss, pred, y, dpred, st, rows, rows * cols);
  SB_LAUNCHED();
}

// fused Softmax + CCE head: p = e/sum(e); loss = -sum(y log(p+eps))/B;
// r_i = y_i/(p_i+eps);  dz_j = -(p_j/B) * (r_j - sum_i r_i p_i)
__global__ void __launch_bounds__(1024) softmax_cce_kernel(float* __restrict__ z, const float* __restrict__ y,
                                                           float* __restrict__ dz, StepState* st, int rows, int C) {
  pdl_prologue();
  const float eps = 1e-8f, fr = (float)rows;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float acc = 0.f;
  for (int row = wid; row < rows; row += nw) {
    float* zr = z + (long long)row * C;
    const float* yr = y + (long long)row * C;
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) {
      float e = expf(zr[c]);
      zr[c] = e;
      sum = __fadd_rn(sum, e);
    }
    sum = warp_sum(sum);
    float dot = 0.f, lsum = 0.f;
    for (int c = lane; c < C; c += 32) {
      float p = __fdiv_rn(zr[c], sum);
      zr[c] = p;
      float pe = __fadd_rn(p, eps);
      lsum = __fadd_rn(lsum, __fmul_rn(yr[c], logf(pe)));
      dot = __fadd_rn(dot, __fmul_rn(__fdiv_rn(yr[c], pe), p));
    }
    dot = warp_sum(dot);
    acc = __fadd_rn(acc, lsum);
    if (dz) {
      float* dr = dz + (long long)row * C;
      for (int c = lane; c < C; c += 32) {
        float p = zr[c];
        float r = __fdiv_rn(yr[c], __fadd_rn(p, eps));
        dr[c] = __fmul_rn(-__fdiv_rn(p, fr), __fsub_rn(r, dot));
      }
    }
  }
  float tot = block_sum_1024(acc);
  if (threadIdx.x == 0) st->loss = __fdiv_rn(-tot, fr);
}
void softmax_cce_fwd_bwd(float* z, const float* y, float* dz, StepState* st, int rows, int C, cudaStream_t s) {
  launch_k(softmax_cce_kernel, 1, 1024, 0, s, z, y, dz, st, rows, C);
  SB_LAUNCHED();
}

// =================================================================================================
// Pool2D (math/nn/AllKernels.scala:270-374).  NHWC, channel fastest => coalesced.
// =================================================================================================
__global__ void pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, PoolGeom g, int reduce, long long n) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int c = (int)(i % g.C);
    long long t = i / g.C;
    int ow = (int)(t % g.OW); t /= g.OW;
    int oh = (int)(t % g.OH);
    int b = (int)(t / g.OH);
    float m = -INFINITY, sum = 0.f;
    int cnt = 0;
    for (int a = 0; a < g.KH; ++a) {
      int ih = oh * g.SH + a - g.PT;
      if (ih < 0 || ih >= g.H) continue;
      for (int bb = 0; bb < g.KW; ++bb) {
        int iw = ow * g.SW + bb - g.PL;
        if (iw < 0 || iw >= g.W) continue;
        float v = __ldg(x + (((long long)b * g.H + ih) * g.W + iw) * g.C + c);
        m = fmaxf(m, v);
        sum = __fadd_rn(sum, v);
        ++cnt;
      }
    }
    y[i] = reduce == 0 ? m : __fdiv_rn(sum, (float)cnt);
  }
}
void pool_fwd(const float* x, float* y, const PoolGeom& g, int reduce, cudaStream_t s) {
  long long n = (long long)g.N * g.OH * g.OW * g.C;
  int blocks = (int)fmin((double)ceil_div(n, 256), (double)kNumSMs * 16);
  launch_k(pool_fwd_kernel, blocks, 256, 0, s, x, y, g, reduce, n);
  SB_LAUNCHED();
}

// gather-form backward: every input element inspects the windows that contain it and takes dy of those
// whose FIRST maximum (row-major (kh,kw) order) is this element.  No atomics, fixed summation order.
template <bool RELU>
__global__ void pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, PoolGeom g,
                                int reduce, float thr, long long n) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int c = (int)(i % g.C);
    long long t = i / g.C;
    int w = (int)(t % g.W); t /= g.W;
    int h = (int)(t % g.H);
    int b = (int)(t / g.H);
    const float* xb = x + (long long)b * g.H * g.W * g.C + c;
    const float self = __ldg(xb + ((long long)h * g.W + w) * g.C);
    float acc = 0.f;
    if (!RELU || self > thr) {
      // windows (oh,ow) with oh*SH - PT <= h < oh*SH - PT + KH
      int oh_lo = h + g.PT - g.KH + 1; oh_lo = oh_lo <= 0 ? 0 : (oh_lo + g.SH - 1) / g.SH;
      int oh_hi = (h + g.PT) / g.SH; if (oh_hi > g.OH - 1) oh_hi = g.OH - 1;
      int ow_lo = w + g.PL - g.KW + 1; ow_lo = ow_lo <= 0 ? 0 : (ow_lo + g.SW - 1) / g.SW;
      int ow_hi = (w + g.PL) / g.SW; if (ow_hi > g.OW - 1) ow_hi = g.OW - 1;
      for (int oh = oh_lo; oh <= oh_hi; ++oh)
        for (int ow = ow_lo; ow <= ow_hi; ++ow) {
          const float gy = __ldg(dy + (((long long)b * g.OH + oh) * g.OW + ow) * g.C + c);
          if (reduce == 0) {
            bool win = true;
            for (int a = 0; a < g.KH && win; ++a) {
              int ih = oh * g.SH + a - g.PT;
              if (ih < 0 || ih >= g.H) continue;
              for (int bb = 0; bb < g.KW; ++bb) {
                int iw = ow * g.SW + bb - g.PL;
                if (iw < 0 || iw >= g.W) continue;
                if (ih == h && iw == w) continue;
                float v = __ldg(xb + ((long long)ih * g.W + iw) * g.C);
                bool earlier = ih < h || (ih == h && iw < w);
                if (earlier ? (v >= self) : (v > self)) { win = false; break; }
              }
            }
            if (win) acc = __fadd_rn(acc, gy);
          } else {
            int cnt = 0;
            for (int a = 0; a < g.KH; ++a) {
              int ih = oh * g.SH + a - g.PT;
              if (ih < 0 || ih >= g.H) continue;
              for (int bb = 0; bb < g.KW; ++bb) {
                int iw = ow * g.SW + bb - g.PL;
                if (iw >= 0 && iw < g.W) ++cnt;
              }
            }
            acc = __fadd_rn(acc, __fdiv_rn(gy, (float)cnt));
          }
        }
    }
    dx[i] = acc;
  }
}
void pool_bwd(const float* x, const float* dy, float* dx, const PoolGeom& g, int reduce, cudaStream_t s) {
  long long n = (long long)g.N * g.H * g.W * g.C;
  int blocks = (int)fmin((double)ceil_div(n, 256), (double)kNumSMs * 16);
  launch_k(pool_bwd_kernel<false>, blocks, 256, 0, s, x, dy, dx, g, reduce, 0.f, n);
  SB_LAUNCHED();
}
void pool_bwd_relu(const float* x, const float* dy, float* dx, const PoolGeom& g, float thr, cudaStream_t s) {
  long long n = (long long)g.N * g.H * g.W * g.C;
  int blocks = (int)fmin((double)ceil_div(n, 256), (double)kNumSMs * 16);
  launch_k(pool_bwd_kernel<true>, blocks, 256, 0, s, x, dy, dx, g, 0, thr, n);
  SB_LAUNCHED();
}

// =================================================================================================
// BatchNorm over the last axis (layer/BatchNorm.scala:107-151; math/nn/AllKernels.scala:399-545,663-684)
// =================================================================================================
struct FBnSum {
  const float* x;
  int C;
  __device__ __forceinline__ float2 operator()(long long r, int c) const { return make_float2(__ldg(x + r * C + c), 0.f); }
};
__global__ void bn_mean_kernel(const float* __restrict__ partial, float* __restrict__ mean, int chunks, int C, float rows) {
  pdl_prologue();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) mean[c] = __fdiv_rn(finalize_partials(partial, chunks, C, c).x, rows);
}
struct FBnSqDev2 {
  const float* x;
  const float* mean;
  int C;
  __device__ __forceinline__ float2 operator()(long long r, int c) const {
    float d = __fsub_rn(__ldg(x + r * C + c), __ldg(mean + c));
    return make_float2(__fmul_rn(d, d), 0.f);
  }
};
// finalises the variance, updates the moving statistics and writes the per-channel scale/shift
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int chunks, int C, float rows, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ mean, float* __restrict__ var_out,
                                   float* __restrict__ mov_mean, float* __restrict__ mov_var, float eps, float momentum,
                                   int fused, int bn_mode, int update_moving) {
  pdl_prologue();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float v = __fdiv_rn(finalize_partials(partial, chunks, C, c).x, rows);  // biased
  var_out[c] = v;
  if (update_moving) {
    float m = mean[c];
    float batch_var;
    if (fused && bn_mode == 0) batch_var = m;  // reference wiring: TakeOut index 1 twice (nn/AllKernels.scala:444-445)
    else if (fused) batch_var = __fmul_rn(v, __fdiv_rn(rows, __fsub_rn(rows, 1.f)));  // TF returns Bessel-corrected
    else batch_var = v;
    float om = __fsub_rn(1.f, momentum);
    mov_mean[c] = __fadd_rn(__fmul_rn(momentum, mov_mean[c]), __fmul_rn(om, m));
    mov_var[c] = __fadd_rn(__fmul_rn(momentum, mov_var[c]), __fmul_rn(om, batch_var));
  }
}
__global__ void bn_apply_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
                                float eps, int fused, long long n, int C) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int c = (int)(i % C);
    float inv = __fmul_rn(__fdiv_rn(1.f, sqrtf(__fadd_rn(__ldg(var + c), eps))), __ldg(gamma + c));
    float m = __ldg(mean + c), xv = x[i];
    float o;
    if (fused) o = __fadd_rn(__fmul_rn(__fsub_rn(xv, m), inv), __ldg(beta + c));
    else o = __fadd_rn(__fsub_rn(__fmul_rn(xv, inv), __fmul_rn(m, inv)), __ldg(beta + c));  // (x*inv - m*inv) + beta
    y[i] = o;
  }
}

// ---- 128-bit BatchNorm kernels (C % 4 == 0): same arithmetic as the scalar functors above, four channels per thread ----
// MODE 0: (sum x, -)   MODE 1: (sum (x - mean)^2, -)   MODE 2: (sum dy, sum dy * (x - mean_in))
template <int MODE>
__global__ void __launch_bounds__(256) bn_col_reduce4_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                             const float* __restrict__ save_mean, const float* __restrict__ save_var,
                                                             float* __restrict__ partial, long long rows, int C, long long rows_per_chunk,
                                                             float frows, int quirk) {
  pdl_prologue();
  __shared__ float4 r0s[8][33];
  __shared__ float4 r1s[8][33];
  const int c4 = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c4 * 4 < C;
  const long long ra = (long long)blockIdx.y * rows_per_chunk;
  const long long rb = ra + rows_per_chunk < rows ? ra + rows_per_chunk : rows;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, mu = a0;
  if (ok && MODE >= 1) {
    mu = *reinterpret_cast<const float4*>(save_mean + c4 * 4);
    if (MODE == 2 && quirk) {  // reference wiring: the "mean" fed back is var_b * R/(R-1)
      const float4 v = *reinterpret_cast<const float4*>(save_var + c4 * 4);
      const float k = __fdiv_rn(frows, __fsub_rn(frows, 1.f));
      mu = make_float4(__fmul_rn(v.x, k), __fmul_rn(v.y, k), __fmul_rn(v.z, k), __fmul_rn(v.w, k));
    }
  }
  if (ok) {
#pragma unroll 4
    for (long long r = ra + threadIdx.y; r < rb; r += 8) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + r * C + c4 * 4));
      if (MODE == 0) {
        a0.x = __fadd_rn(a0.x, xv.x); a0.y = __fadd_rn(a0.y, xv.y); a0.z = __fadd_rn(a0.z, xv.z); a0.w = __fadd_rn(a0.w, xv.w);
      } else if (MODE == 1) {
        const float dx_ = __fsub_rn(xv.x, mu.x), dy_ = __fsub_rn(xv.y, mu.y), dz_ = __fsub_rn(xv.z, mu.z), dw_ = __fsub_rn(xv.w, mu.w);
        a0.x = __fadd_rn(a0.x, __fmul_rn(dx_, dx_)); a0.y = __fadd_rn(a0.y, __fmul_rn(dy_, dy_));
        a0.z = __fadd_rn(a0.z, __fmul_rn(dz_, dz_)); a0.w = __fadd_rn(a0.w, __fmul_rn(dw_, dw_));
      } else {
        const float4 g = __ldg(reinterpret_cast<const float4*>(dy + r * C + c4 * 4));
        a0.x = __fadd_rn(a0.x, g.x); a0.y = __fadd_rn(a0.y, g.y); a0.z = __fadd_rn(a0.z, g.z); a0.w = __fadd_rn(a0.w, g.w);
        a1.x = __fadd_rn(a1.x, __fmul_rn(g.x, __fsub_rn(xv.x, mu.x))); a1.y = __fadd_rn(a1.y, __fmul_rn(g.y, __fsub_rn(xv.y, mu.y)));
        a1.z = __fadd_rn(a1.z, __fmul_rn(g.z, __fsub_rn(xv.z, mu.z))); a1.w = __fadd_rn(a1.w, __fmul_rn(g.w, __fsub_rn(xv.w, mu.w)));
      }
    }
  }
  r0s[threadIdx.y][threadIdx.x] = a0;
  r1s[threadIdx.y][threadIdx.x] = a1;
  __syncthreads();
  if (threadIdx.y == 0 && ok) {
    for (int j = 1; j < 8; ++j) {
      const float4 u = r0s[j][threadIdx.x], v = r1s[j][threadIdx.x];
      a0.x = __fadd_rn(a0.x, u.x); a0.y = __fadd_rn(a0.y, u.y); a0.z = __fadd_rn(a0.z, u.z); a0.w = __fadd_rn(a0.w, u.w);
      a1.x = __fadd_rn(a1.x, v.x); a1.y = __fadd_rn(a1.y, v.y); a1.z = __fadd_rn(a1.z, v.z); a1.w = __fadd_rn(a1.w, v.w);
    }
    // same partial layout as col_reduce_kernel: [chunk][2][C]
    *reinterpret_cast<float4*>(partial + ((long long)blockIdx.y * 2 + 0) * C + c4 * 4) = a0;
    *reinterpret_cast<float4*>(partial + ((long long)blockIdx.y * 2 + 1) * C + c4 * 4) = a1;
  }
}
template <int MODE>
static ColChunks launch_bn_reduce4(const float* x, const float* dy, const float* save_mean, const float* save_var, float* ws, size_t wsf,
                                   long long rows, int C, int quirk, cudaStream_t s) {
  const int col_blocks = ceil_div(C / 4, 32);
  long long want = ceil_div(4 * kNumSMs, col_blocks);
  if (want > (rows + 31) / 32) want = (rows + 31) / 32;
  const long long max_by_ws = (long long)(wsf / ((size_t)2 * C));
  if (want > max_by_ws) want = max_by_ws;
  if (want < 1) want = 1;
  const long long rpc = (rows + want - 1) / want;
  const int chunks = ceil_div(rows, rpc);
  launch_k(bn_col_reduce4_kernel<MODE>, dim3(col_blocks, chunks), dim3(32, 8), 0, s, x, dy, save_mean, save_var, ws, rows, C, rpc, (float)rows, quirk);
  SB_LAUNCHED();
  return ColChunks{chunks, rpc};
}
__global__ void __launch_bounds__(256) bn_apply4_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ mean,
                                                        const float* __restrict__ var, float eps, int fused, long long n4, int C4) {
  pdl_prologue();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int c = (int)(i % C4) * 4;
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 vr = __ldg(reinterpret_cast<const float4*>(var + c)), gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 mn = __ldg(reinterpret_cast<const float4*>(mean + c)), bt = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 o;
#define SB_BN1(f)                                                                                      \
    {                                                                                                  \
      const float inv = __fmul_rn(__fdiv_rn(1.f, sqrtf(__fadd_rn(vr.f, eps))), gm.f);                \
      o.f = fused ? __fadd_rn(__fmul_rn(__fsub_rn(xv.f, mn.f), inv), bt.f)                            \
                  : __fadd_rn(__fsub_rn(__fmul_rn(xv.f, inv), __fmul_rn(mn.f, inv)), bt.f);           \
    }
    SB_BN1(x) SB_BN1(y) SB_BN1(z) SB_BN1(w)
#undef SB_BN1
    reinterpret_cast<float4*>(y)[i] = o;
  }
}
__global__ void __launch_bounds__(256) bn_bwd_apply4_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                                            const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                                            const float* __restrict__ save_var, const float* __restrict__ sums, float eps,
                                                            float rows, int quirk, long long n4, int C) {
  pdl_prologue();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int C4 = C >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int c = (int)(i % C4) * 4;
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i), g = __ldg(reinterpret_cast<const float4*>(dy) + i);
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(save_mean + c)), v4 = __ldg(reinterpret_cast<const float4*>(save_var + c));
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(sums + c)), s1 = __ldg(reinterpret_cast<const float4*>(sums + C + c));
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
    float4 o;
#define SB_BNB(f)                                                                                                   \
    {                                                                                                               \
      const float mean_in = quirk ? __fmul_rn(v4.f, __fdiv_rn(rows, __fsub_rn(rows, 1.f))) : m4.f;                 \
      const float ve = __fadd_rn(quirk ? m4.f : v4.f, eps);                                                        \
      const float istd = __fdiv_rn(1.f, sqrtf(ve));                                                                \
      const float mg = __fdiv_rn(s0.f, rows), mgx = __fdiv_rn(s1.f, rows);                                         \
      const float xc = __fsub_rn(xv.f, mean_in);                                                                   \
      const float t = __fsub_rn(__fsub_rn(g.f, mg), __fdiv_rn(__fmul_rn(xc, mgx), ve));                            \
      o.f = __fmul_rn(__fmul_rn(gm.f, istd), t);                                                                   \
    }
    SB_BNB(x) SB_BNB(y) SB_BNB(z) SB_BNB(w)
#undef SB_BNB
    reinterpret_cast<float4*>(dx)[i] = o;
  }
}

void bn_fwd(const float* x, float* y, const float* gamma, const float* beta, float* mov_mean, float* mov_var,
            float* save_mean, float* save_var, long long rows, int C, float eps, float momentum, int fused, int bn_mode,
            int update_moving, float* ws, size_t wsf, cudaStream_t s) {
  const bool v4 = (C % 4 == 0) && ((((uintptr_t)x) | ((uintptr_t)y)) % 16 == 0);
  const char* two_pass = getenv("SB_BN_TWO_PASS");
  const bool lean = !(two_pass && two_pass[0] == '1');
  // one pass over x (bn_pool.cu) where the layout allows; else the two-pass moments: mean, then mean of squared deviations
  if (!(lean && v4 && bn_stats(x, rows, C, save_mean, save_var, mov_mean, mov_var, momentum, fused, bn_mode, update_moving, ws, wsf, s))) {
    ColChunks cc = v4 ? launch_bn_reduce4<0>(x, nullptr, nullptr, nullptr, ws, wsf, rows, C, 0, s)
                      : launch_col_reduce(FBnSum{x, C}, ws, wsf, rows, C, s);
    launch_k(bn_mean_kernel, ceil_div(C, 128), 128, 0, s, ws, save_mean, cc.chunks, C, (float)rows);
    SB_LAUNCHED();
    cc = v4 ? launch_bn_reduce4<1>(x, nullptr, save_mean, nullptr, ws, wsf, rows, C, 0, s)
            : launch_col_reduce(FBnSqDev2{x, save_mean, C}, ws, wsf, rows, C, s);
    launch_k(bn_finalize_kernel, ceil_div(C, 128), 128, 0, s, ws, cc.chunks, C, (float)rows, gamma, beta, save_mean, save_var, mov_mean,
                                                        mov_var, eps, momentum, fused, bn_mode, update_moving);
    SB_LAUNCHED();
  }
  long long n = rows * C;
  if (v4) {
    int blocks = (int)fmin((double)ceil_div(n / 4, 256), (double)kNumSMs * 16);
    launch_k(bn_apply4_kernel, blocks, 256, 0, s, x, y, gamma, beta, save_mean, save_var, eps, fused, n / 4, C / 4);
    SB_LAUNCHED();
    return;
  }
  int blocks = (int)fmin((double)ceil_div(n, 256), (double)kNumSMs * 16);
  launch_k(bn_apply_kernel, blocks, 256, 0, s, x, y, gamma, beta, save_mean, save_var, eps, fused, n, C);
  SB_LAUNCHED();
}

void bn_fwd_infer(const float* x, float* y, const float* gamma, const float* beta, const float* mov_mean, const float* mov_var,
                  long long rows, int C, float eps, int fused, cudaStream_t s) {
  // `freeze` => is_training=false: normalise with the moving statistics (layer/BatchNorm.scala:126)
  long long n = rows * C;
  int blocks = (int)fmin((double)ceil_div(n, 256), (double)kNumSMs * 16);
  launch_k(bn_apply_kernel, blocks, 256, 0, s, x, y, gamma, beta, mov_mean, mov_var, eps, fused, n, C);
  SB_LAUNCHED();
}

// backward: sums dbeta = sum(dy), dgamma = sum(dy*xc*istd); dx = gamma*istd*(dy - mean(dy) - xc*mean(dy*xc)/(var_in+eps))
// reference mode (fused): mean_in = var_b*R/(R-1), var_in = batch mean
struct FBnBwd {
  const float* x;
  const float* dy;
  const float* save_mean;
  const float* save_var;
  int C;
  float rows;
  int quirk;
  __device__ __forceinline__ float2 operator()(long long r, int c) const {
    float m = __ldg(save_mean + c), v = __ldg(save_var + c);
    float mean_in = quirk ? __fmul_rn(v, __fdiv_rn(rows, __fsub_rn(rows, 1.f))) : m;
    float g = __ldg(dy + r * C + c);
    float xc = __fsub_rn(__ldg(x + r * C + c), mean_in);
    return make_float2(g, __fmul_rn(g, xc));
  }
};
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int chunks, int C, const float* __restrict__ save_mean,
                                       const float* __restrict__ save_var, float eps, int quirk, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ sums) {
  pdl_prologue();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float2 sv = finalize_partials(partial, chunks, C, c);
  float var_in = quirk ? save_mean[c] : save_var[c];
  float istd = __fdiv_rn(1.f, sqrtf(__fadd_rn(var_in, eps)));
  dbeta[c] = sv.x;
  dgamma[c] = __fmul_rn(sv.y, istd);
  sums[c] = sv.x;
  sums[C + c] = sv.y;
}
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                    const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                    const float* __restrict__ save_var, const float* __restrict__ sums, float eps, float rows,
                                    int quirk, long long n, int C) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int c = (int)(i % C);
    float m = __ldg(save_mean + c), v = __ldg(save_var + c);
    float mean_in = quirk ? __fmul_rn(v, __fdiv_rn(rows, __fsub_rn(rows, 1.f))) : m;
    float var_in = quirk ? m : v;
    float ve = __fadd_rn(var_in, eps);
    float istd = __fdiv_rn(1.f, sqrtf(ve));
    float mg = __fdiv_rn(__ldg(sums + c), rows), mgx = __fdiv_rn(__ldg(sums + C + c), rows);
    float xc = __fsub_rn(x[i], mean_in);
    float t = __fsub_rn(__fsub_rn(dy[i], mg), __fdiv_rn(__fmul_rn(xc, mgx), ve));
    dx[i] = __fmul_rn(__fmul_rn(__ldg(gamma + c), istd), t);
  }
}
void bn_bwd(const float* x, const float* dy, float* dx, const float* gamma, const float* save_mean, const float* save_var,
            float* dgamma, float* dbeta, long long rows, int C, float eps, int fused, int bn_mode, float* ws, size_t wsf,
            cudaStream_t s) {
  const int quirk = (fused && bn_mode == 0) ? 1 : 0;
  // ws layout: [2*C sums][partials...]
  float* sums = ws;
  float* part = ws + 2 * (size_t)C;
  const bool v4 = (C % 4 == 0) && ((((uintptr_t)x) | ((uintptr_t)dy) | ((uintptr_t)dx)) % 16 == 0);
  ColChunks cc = v4 ? launch_bn_reduce4<2>(x, dy, save_mean, save_var, part, wsf - 2 * (size_t)C, rows, C, quirk, s)
                    : launch_col_reduce(FBnBwd{x, dy, save_mean, save_var, C, (float)rows, quirk}, part, wsf - 2 * (size_t)C, rows, C, s);
  launch_k(bn_bwd_finalize_kernel, ceil_div(C, 128), 128, 0, s, part, cc.chunks, C, save_mean, save_var, eps, quirk, dgamma, dbeta, sums);
  SB_LAUNCHED();
  if (dx && v4) {
    const long long n4 = rows * C / 4;
    int blocks = (int)fmin((double)ceil_div(n4, 256), (double)kNumSMs * 16);
    launch_k(bn_bwd_apply4_kernel, blocks, 256, 0, s, x, dy, dx, gamma, save_mean, save_var, sums, eps, (float)rows, quirk, n4, C);
    SB_LAUNCHED();
  } else if (dx) {
    long long n = rows * C;
    int blocks = (int)fmin((double)ceil_div(n, 256), (double)kNumSMs * 16);
    launch_k(bn_bwd_apply_kernel, blocks, 256, 0, s, x, dy, dx, gamma, save_mean, save_var, sums, eps, (float)rows, quirk, n, C);
    SB_LAUNCHED();
  }
}

// =================================================================================================
// regularisation penalties (models/Regularization.scala:16-44): one CTA per tensor, deterministic tree
// =================================================================================================
__global__ void __launch_bounds__(1024) reg_penalty_kernel(const float* __restrict__ w, float* __restrict__ g, long long n, int kind,
                                                           float lambda, StepState* st, int first, int last) {
  pdl_prologue();
  float acc = 0.f;
  const float gl1 = __fdiv_rn(fabsf(lambda), 2.f);
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float wi = w[i];
    acc = __fadd_rn(acc, kind == 2 ? __fmul_rn(wi, wi) : fabsf(wi));
    if (g) g[i] = __fadd_rn(g[i], kind == 2 ? __fmul_rn(lambda, wi) : gl1);
  }
  const float tot = block_sum_1024(acc);
  if (threadIdx.x == 0) {
    const float pen = __fdiv_rn(__fmul_rn(lambda, tot), 2.f);
    const float run = first ? pen : __fadd_rn(st->pad, pen);  // penalties add up in tensor order ...
    st->pad = run;
    if (last) st->loss = __fadd_rn(st->loss, run);            // ... and the total joins the loss (Model.scala:97)
  }
}
void reg_penalty_grad(const float* w, float* g, long long n, int kind, float lambda, StepState* st, int first, int last,
                      cudaStream_t s) {
  launch_k(reg_penalty_kernel, 1, 1024, 0, s, w, g, n, kind, lambda, st, first, last);
  SB_LAUNCHED();
}

// =================================================================================================
// estimators (estimators/package.scala:18-137): per-batch partial sums, accumulated in double on the device
// =================================================================================================
__global__ void __launch_bounds__(256) estimator_kernel(const float* __restrict__ pred, const float* __restrict__ y, long long rows,
                                                        int C, int kind, double* __restrict__ acc) {
  pdl_prologue();
  double a0 = 0.0, a2 = 0.0, a3 = 0.0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const float* p = pred + r * C;
    const float* t = y + r * C;
    if (kind == 0) {  // yPredicted = result.round ; positive when every output matches (package.scala:30-33)
      bool all = true;
      for (int c = 0; c < C; ++c) all = all && (rintf(p[c]) == t[c]);
      a0 += all ? 1.0 : 0.0;
    } else {
      for (int c = 0; c < C; ++c) {
        const float d = __fsub_rn(p[c], t[c]);
        a0 += kind == 2 ? (double)fabsf(d) : (double)__fmul_rn(d, d);
        if (kind == 3) { a2 += (double)p[c]; a3 += (double)p[c] * (double)p[c]; }
      }
    }
  }
  __shared__ double sh[3][8];
  for (int o = 16; o; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = a0; sh[1][w] = a2; sh[2][w] = a3; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0, s2 = 0, s3 = 0;
    for (int i = 0; i < 8; ++i) { s0 += sh[0][i]; s2 += sh[1][i]; s3 += sh[2][i]; }
    atomicAdd(acc + 0, s0);
    if (blockIdx.x == 0) atomicAdd(acc + 1, (double)rows);
    if (kind == 3) { atomicAdd(acc + 2, s2); atomicAdd(acc + 3, s3); }
  }
}
void estimator_accumulate(const float* pred, const float* y, long long rows, int C, int kind, double* acc, cudaStream_t s) {
  if (rows <= 0) return;
  const int grid = (int)std::min<long long>((rows + 255) / 256, 148 * 4);
  launch_k(estimator_kernel, grid, 256, 0, s, pred, y, rows, C, kind, acc);
  SB_LAUNCHED();
}

// =================================================================================================
// fused multi-tensor optimizer over the flat arena (optimizers/*.scala; caller does w - delta,
// Optimizer.scala:186).  28 B/param for Adam: reads w,g,m,v ; writes w,m,v.
// =================================================================================================
struct OptScalars {
  float rate, b1, b2, eps, om_b1, om_b2, bc1, bc2;
  int nesterov;
  int maximize;  // Optimizer.maximize: w + delta instead of w - delta (optimizers/Optimizer.scala:186)
};
__device__ __forceinline__ float decaying_avg(float avg, float nxt, float d, float om_d) {
  return __fadd_rn(__fmul_rn(d, avg), __fmul_rn(om_d, nxt));  // (d*avg) + ((1-d)*next), alg:541-544
}
template <int KIND>
__device__ __forceinline__ void opt_update(const OptScalars& o, float& w, float g, float& s0, float& s1, float& s2) {
  float delta;
  if (KIND == 0) {  // SGD.scala:15-25
    float rg = __fmul_rn(o.rate, g);
    float v = __fsub_rn(rg, __fmul_rn(o.b1, s0));
    delta = o.nesterov ? __fsub_rn(rg, __fmul_rn(o.b1, v)) : v;
    s0 = v;
  } else if (KIND == 1) {  // Adam.scala:29-41: (rate / sqrt(boost(v) + eps)) * boost(m)
    float m = decaying_avg(s0, g, o.b1, o.om_b1);
    float v = decaying_avg(s1, __fmul_rn(g, g), o.b2, o.om_b2);
    float vb = __fdiv_rn(v, o.bc2), mb = __fdiv_rn(m, o.bc1);
    delta = __fmul_rn(__fdiv_rn(o.rate, sqrtf(__fadd_rn(vb, o.eps))), mb);
    s0 = m; s1 = v;
  } else if (KIND == 2) {  // AdaGrad.scala:21-32
    float a = __fadd_rn(s0, __fmul_rn(g, g));
    delta = __fmul_rn(__fdiv_rn(o.rate, sqrtf(__fadd_rn(a, o.eps))), g);
    s0 = a;
  } else if (KIND == 3) {  // RMSProp.scala:24-32 (rho in b1)
    float a = decaying_avg(s0, __fmul_rn(g, g), o.b1, o.om_b1);
    delta = __fmul_rn(__fdiv_rn(o.rate, sqrtf(__fadd_rn(a, o.eps))), g);
    s0 = a;
  } else if (KIND == 4) {  // AdaDelta.scala:33-44 (rho in b1): s0 avg_grad, s1 avg_delta
    float a = decaying_avg(s0, __fmul_rn(g, g), o.b1, o.om_b1);
    float ratio = __fdiv_rn(sqrtf(__fadd_rn(s1, o.eps)), sqrtf(__fadd_rn(a, o.eps)));
    delta = __fmul_rn(o.rate, __fmul_rn(ratio, g));
    s1 = decaying_avg(s1, __fmul_rn(delta, delta), o.b1, o.om_b1);
    s0 = a;
  } else if (KIND == 5) {  // Adamax.scala:30-52: (rate * boost(m1)) / (m2 + eps)
    float m1 = decaying_avg(s0, g, o.b1, o.om_b1);
    float m2 = fmaxf(__fmul_rn(o.b2, s1), __fmul_rn(o.om_b2, fabsf(g)));
    delta = __fdiv_rn(__fmul_rn(o.rate, __fdiv_rn(m1, o.bc1)), __fadd_rn(m2, o.eps));
    s0 = m1; s1 = m2;
  } else if (KIND == 6) {  // Nadam.scala:26-49 (no learning rate)
    float m = decaying_avg(s0, g, o.b1, o.om_b1);
    float v = decaying_avg(s1, __fmul_rn(g, g), o.b2, o.om_b2);
    float mn = __fadd_rn(__fmul_rn(o.b1, __fdiv_rn(m, o.bc1)), __fmul_rn(o.om_b1, __fdiv_rn(g, o.bc1)));
    delta = __fdiv_rn(mn, __fadd_rn(sqrtf(__fdiv_rn(v, o.bc2)), o.eps));
    s0 = m; s1 = v;
  } else {  // AMSGrad.scala:34-46: (rate*m) / (sqrt(vhat) + eps), no bias correction
    float m = decaying_avg(s0, g, o.b1, o.om_b1);
    float vc = decaying_avg(s1, __fmul_rn(g, g), o.b2, o.om_b2);
    float vh = fmaxf(s2, vc);
    delta = __fdiv_rn(__fmul_rn(o.rate, m), __fadd_rn(sqrtf(vh), o.eps));
    s0 = m; s1 = vc; s2 = vh;
  }
  w = o.maximize ? __fadd_rn(w, delta) : __fsub_rn(w, delta);  // `if (minimizing) w - del else w + del` (Optimizer.scala:186)
}
template <int KIND>
__global__ void __launch_bounds__(256) optimizer_kernel(OptDesc od, float4* __restrict__ w, const float4* __restrict__ g,
                                                        float4* __restrict__ s0, float4* __restrict__ s1,
                                                        float4* __restrict__ s2, long long n4, StepState* st,
                                                        int advance_batch, int early) {
  pdl_prologue_trigger_if(early);
  OptScalars o;
  o.rate = od.rate; o.b1 = od.beta1; o.b2 = od.beta2; o.eps = od.epsilon; o.nesterov = od.nesterov;
  o.maximize = od.minimizing ? 0 : 1;
  o.om_b1 = __fsub_rn(1.f, od.beta1);
  o.om_b2 = __fsub_rn(1.f, od.beta2);
  o.bc1 = st->bc1; o.bc2 = st->bc2;
  constexpr int NS = (KIND == 7) ? 3 : ((KIND == 1 || KIND == 4 || KIND == 5 || KIND == 6) ? 2 : 1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 wv = w[i];
    const float4 gv = g[i];
    float4 a = s0[i];
    float4 b = NS >= 2 ? s1[i] : make_float4(0, 0, 0, 0);
    float4 c = NS >= 3 ? s2[i] : make_float4(0, 0, 0, 0);
    opt_update<KIND>(o, wv.x, gv.x, a.x, b.x, c.x);
    opt_update<KIND>(o, wv.y, gv.y, a.y, b.y, c.y);
    opt_update<KIND>(o, wv.z, gv.z, a.z, b.z, c.z);
    opt_update<KIND>(o, wv.w, gv.w, a.w, b.w, c.w);
    w[i] = wv;
    s0[i] = a;
    if (NS >= 2) s1[i] = b;
    if (NS >= 3) s2[i] = c;
  }
  // the batch cursor is only read by the NEXT step's prologue kernel, so it can be advanced here
  if (advance_batch && blockIdx.x == 0 && threadIdx.x == 0) st->batch_idx += 1;
  // the NEXT step's bias corrections 1 - beta^t (`boost`, math/alg/AllKernels.scala:546-549): the double-precision pow pair is the
  // longest dependent chain of the prologue kernel; here it hides behind the update of the arena.  bc1 / bc2 themselves are left
  // alone (other threads of this grid are still reading them); the prologue adopts the pair if `iter` is still this step's.
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == blockDim.x - 1) {
    const long long t = st->iter + 1;
    const double tf = (double)(float)t;
    st->bc1_next = __fsub_rn(1.0f, (float)pow((double)od.beta1, tf));
    st->bc2_next = __fsub_rn(1.0f, (float)pow((double)od.beta2, tf));
    st->bc_iter = t;
  }
}
void optimizer_step(const OptDesc& o, float* w, const float* g, float* s0, float* s1, float* s2, long long n, StepState* st,
                    int advance_batch, cudaStream_t s) {
  long long n4 = n / 4;  // the arena pads every tensor to 32 floats
  int blocks = (int)fmin((double)ceil_div(n4, 256), (double)kNumSMs * 8);
  if (blocks < 1) blocks = 1;
#define SB_OPT(K)                                                                                             \
  case K:                                                                                                     \
    launch_k(optimizer_kernel<K>, blocks, 256, 0, s, o, (float4*)w, (const float4*)g, (float4*)s0, (float4*)s1, \
                                                 (float4*)s2, n4, st, advance_batch, pdl_early_hint());        \
    break;
  switch (o.kind) {
    SB_OPT(0) SB_OPT(1) SB_OPT(2) SB_OPT(3) SB_OPT(4) SB_OPT(5) SB_OPT(6) SB_OPT(7)
    default: throw std::runtime_error("unknown optimizer kind");
  }
#undef SB_OPT
  SB_LAUNCHED();
  set_prewait_ok(0);  // writes weights / optimizer state: the next kernel must not read weights before its wait
}

// =================================================================================================
// init / utility
// =================================================================================================
__global__ void fill_kernel(float* __restrict__ p, float v, long long n) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}
// Busy-wait `ns` nanoseconds on one thread.  Profiling mode enqueues it ahead of a batch of event-bracketed launches so the
// host runs ahead of the device and the events measure kernel time, not launch gaps.
__global__ void delay_kernel(long long ns) {
  pdl_prologue();
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    __nanosleep(1000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while ((long long)(t - t0) < ns);
}
void device_delay(long long ns, cudaStream_t s) {
  launch_k(delay_kernel, 1, 1, 0, s, ns);
  SB_CUDA(cudaPeekAtLastError());
}
void fill_const(float* p, float v, long long n, cudaStream_t s) {
  if (n <= 0) return;
  int blocks = (int)fmin((double)ceil_div(n, 256), (double)kNumSMs * 16);
  launch_k(fill_kernel, blocks, 256, 0, s, p, v, n);
  SB_LAUNCHED();
  set_prewait_ok(0);  // writes weights / optimizer state: the next kernel must not read weights before its wait
}
__global__ void scale_kernel(float* __restrict__ p, float w, long long n) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = __fmul_rn(p[i], w);
}
void scale_inplace(float* p, float w, long long n, cudaStream_t s) {
  if (n <= 0) return;
  int blocks = (int)fmin((double)ceil_div(n, 256), (double)kNumSMs * 16);
  launch_k(scale_kernel, blocks, 256, 0, s, p, w, n);
  SB_LAUNCHED();
  set_prewait_ok(0);  // writes weights / optimizer state: the next kernel must not read weights before its wait
}

// TF's seeded RandomUniform / RandomStandardNormal: Philox-4x32-10, key = seed, counter = block index
// (math/rand/AllKernels.scala:12-43 set only `seed`; seed2 = 0).  Bit-compatible with oracle/philox.py.
__device__ __forceinline__ void philox_block(unsigned long long seed, unsigned long long ctr, unsigned int out[4]) {
  unsigned int c0 = (unsigned int)ctr, c1 = (unsigned int)(ctr >> 32), c2 = 0u, c3 = 0u;
  unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float u32_to_float(unsigned int x) {
  return __fsub_rn(__uint_as_float((127u << 23) | (x & 0x7fffffu)), 1.0f);
}
__global__ void philox_fill_kernel(float* __restrict__ p, long long n, unsigned long long seed, int normal, float scale,
                                   int has_scale, float shift, int has_shift) {
  pdl_prologue();
  long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nblk = (n + 3) / 4;
  if (blk >= nblk) return;
  unsigned int w[4];
  philox_block(seed, (unsigned long long)blk, w);
  float v[4];
  if (normal == 2) {
    // TF TruncatedNormal(float32): every group of 4 outputs restarts the generator at counter group*256 and draws word pairs
    // (Box-Muller) until four values with |x| < 2 are found (random_distributions.h TruncatedNormalDistribution)
    int got = 0;
    unsigned long long ctr = (unsigned long long)blk * 256ull;
    while (got < 4) {
      philox_block(seed, ctr++, w);
      for (int j = 0; j < 4 && got < 4; j += 2) {
        float u1 = fmaxf(u32_to_float(w[j]), 1.0e-7f);
        float v1 = (float)(6.283185307179586 * (double)u32_to_float(w[j + 1]));
        float u2 = sqrtf(__fmul_rn(-2.0f, logf(u1)));
        float sn, cs;
        sincosf(v1, &sn, &cs);
        const float f0 = __fmul_rn(sn, u2), f1 = __fmul_rn(cs, u2);
        if (fabsf(f0) < 2.0f) v[got++] = f0;
        if (got < 4 && fabsf(f1) < 2.0f) v[got++] = f1;
      }
    }
  } else if (!normal) {
    for (int j = 0; j < 4; ++j) v[j] = u32_to_float(w[j]);
  } else {
    for (int j = 0; j < 4; j += 2) {  // BoxMullerFloat
      float u1 = fmaxf(u32_to_float(w[j]), 1.0e-7f);
      float v1 = (float)(6.283185307179586 * (double)u32_to_float(w[j + 1]));  // `2.0f * M_PI * u` is a double product
      float u2 = sqrtf(__fmul_rn(-2.0f, logf(u1)));
      float sn, cs;
      sincosf(v1, &sn, &cs);
      v[j] = __fmul_rn(sn, u2);
      v[j + 1] = __fmul_rn(cs, u2);
    }
  }
  for (int j = 0; j < 4; ++j) {
    long long i = blk * 4 + j;
    if (i < n) {
      float o = v[j];
      if (has_scale) o = __fmul_rn(o, scale);
      if (has_shift) o = __fadd_rn(o, shift);
      p[i] = o;
    }
  }
}
void philox_fill(float* p, long long n, unsigned long long seed, int normal, float scale, int has_scale, float shift,
                 int has_shift, cudaStream_t s) {
  if (n <= 0) return;
  long long nblk = (n + 3) / 4;
  launch_k(philox_fill_kernel, ceil_div(nblk, 256), 256, 0, s, p, n, seed, normal, scale, has_scale, shift, has_shift);
  SB_LAUNCHED();
  set_prewait_ok(0);  // writes weights / optimizer state: the next kernel must not read weights before its wait
}

}  // namespace sb

namespace sb {
void trace_set_kernels_simt(unsigned long long* p) { sb_trace_set_local(p); }  // SB_TRACE timeline buffer of this translation unit
}  // namespace sb
