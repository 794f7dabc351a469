// BatchNorm on NHWC activations, bandwidth-lean forms (layer/BatchNorm.scala:107-151; math/nn/AllKernels.scala:399-545,663-684):
//   * bn_stats           -- batch mean and (biased) variance in ONE pass over x: every thread keeps a running (n, mean, M2) per
//                           channel and folds eight rows at a time into it (Chan et al.'s pairwise update, the numerically stable
//                           form of the two-pass moments the reference computes); threads, then chunks, are merged in a fixed order.
//   * bn_apply_pool_fwd  -- BatchNorm >> Pool2D(2x2, stride 2, VALID, max): the normalised activations are never written; the
//                           kernel stores the pooled tensor and one winner byte per pooled element (win4's first-max rule).
//   * bn_pool_bwd        -- the backward of that pair (+ the in-place ReLU in front of the BatchNorm): the pooled gradient is routed
//                           by the winner bytes while the column sums (dbeta, dgamma) and dx are formed, so neither the pool's dx
//                           nor the ReLU's ever goes to memory.  x is read twice, dx written once.
// Layout: channel-fastest, C % 4 == 0 and 256 % (C/4) == 0, so that with 256 threads per block a thread always sees the same four
// channels (thread t: channel group t % (C/4), row t / (C/4) of every block-row of 256/(C/4) rows) and every access is one
// fully-used 128-bit load.  All reductions use fixed orders (no atomics): results are run-to-run deterministic.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace sb {
namespace {

constexpr int kBnThreads = 256;
constexpr int kBnUnroll = 8;   // rows folded per update (and loads in flight per thread)
constexpr int kBnSegments = 32; // the finalize kernels sum the chunks in 32 interleaved segments, then the segments in order

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// running moments (n, mean m, M2 s) <- merged with a batch (nb, mb, sb); n == 0 is fine (the batch is taken as is)
__device__ __forceinline__ void chan_merge(float& n, float& m, float& s, float nb, float mb, float sb) {
  const float tot = __fadd_rn(n, nb);
  const float f = __fdiv_rn(nb, tot);
  const float d = __fsub_rn(mb, m);
  s = __fadd_rn(__fadd_rn(s, sb), __fmul_rn(__fmul_rn(d, d), __fmul_rn(n, f)));
  m = __fadd_rn(m, __fmul_rn(d, f));
  n = tot;
}

// partial[chunk][2][C]: (mean, M2) of the chunk's rows; the row count of a chunk follows from its index
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* __restrict__ x, float* __restrict__ partial, long long rows, int C4,
                                                             long long rows_per_chunk) {
  pdl_prologue();
  __shared__ float4 sm_m[kBnThreads];
  __shared__ float4 sm_s[kBnThreads];
  __shared__ float sm_n[kBnThreads];
  const int t = threadIdx.x, c4 = t % C4, j = t / C4, R = kBnThreads / C4;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  float m[4] = {0.f, 0.f, 0.f, 0.f}, s[4] = {0.f, 0.f, 0.f, 0.f};
  float n = 0.f;
  const float4* xp = reinterpret_cast<const float4*>(x) + c4;
  long long r = r0 + j;
  for (; r + (long long)(kBnUnroll - 1) * R < r1; r += (long long)kBnUnroll * R) {
    float v[kBnUnroll][4];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const float4 q = __ldg(xp + (r + (long long)u * R) * C4);
      v[u][0] = q.x; v[u][1] = q.y; v[u][2] = q.z; v[u][3] = q.w;
    }
    const float tot = n + (float)kBnUnroll;
    const float f = __fdiv_rn((float)kBnUnroll, tot), nf = __fmul_rn(n, f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float sum = v[0][k];
#pragma unroll
      for (int u = 1; u < kBnUnroll; ++u) sum = __fadd_rn(sum, v[u][k]);
      const float bm = __fmul_rn(sum, 1.f / kBnUnroll);
      float bs = 0.f;
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const float d = __fsub_rn(v[u][k], bm);
        bs = __fmaf_rn(d, d, bs);
      }
      const float d = __fsub_rn(bm, m[k]);
      s[k] = __fadd_rn(__fadd_rn(s[k], bs), __fmul_rn(__fmul_rn(d, d), nf));
      m[k] = __fadd_rn(m[k], __fmul_rn(d, f));
    }
    n = tot;
  }
  for (; r < r1; r += R) {  // the chunk's last rows, one at a time
    const float4 q = __ldg(xp + r * C4);
    const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float nn = n;
      chan_merge(nn, m[k], s[k], 1.f, v[k], 0.f);
    }
    n += 1.f;
  }
  sm_m[t] = make_float4(m[0], m[1], m[2], m[3]);
  sm_s[t] = make_float4(s[0], s[1], s[2], s[3]);
  sm_n[t] = n;
  __syncthreads();
  if (j == 0) {
    for (int jj = 1; jj < R; ++jj) {
      const float nb = sm_n[jj * C4 + c4];
      if (nb == 0.f) continue;
      const float4 mb = sm_m[jj * C4 + c4], sb = sm_s[jj * C4 + c4];
      float n0 = n, n1 = n, n2 = n, n3 = n;
      chan_merge(n0, m[0], s[0], nb, mb.x, sb.x);
      chan_merge(n1, m[1], s[1], nb, mb.y, sb.y);
      chan_merge(n2, m[2], s[2], nb, mb.z, sb.z);
      chan_merge(n3, m[3], s[3], nb, mb.w, sb.w);
      n = n0;
    }
    const int C = C4 * 4;
    *reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * 2 + 0) * C + c4 * 4) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * 2 + 1) * C + c4 * 4) = make_float4(s[0], s[1], s[2], s[3]);
  }
}

// merges the chunk moments and writes the batch statistics; updates the moving ones exactly as bn_finalize_kernel (kernels_simt.cu).
// With n_j rows, mean m_j and M2_j per chunk:  mean = sum n_j m_j / N,  M2 = sum (M2_j + n_j (m_j - mean)^2)  (the pairwise-merge
// identity summed over all chunks at once, so no serial chain of divisions).  Thread (channel, segment): chunks segment, segment + 32, ..
// in ascending order, then the 32 segments in order.
__global__ void __launch_bounds__(32 * kBnSegments) bn_stats_finalize_kernel(const float* __restrict__ partial, int chunks, int C, long long rows,
                                                                             long long rows_per_chunk, float* __restrict__ mean_out,
                                                                             float* __restrict__ var_out, float* __restrict__ mov_mean,
                                                                             float* __restrict__ mov_var, float momentum, int fused, int bn_mode,
                                                                             int update_moving) {
  pdl_prologue();
  __shared__ float red[kBnSegments][33];
  __shared__ float mean_s[32];
  const int c = blockIdx.x * 32 + threadIdx.x, seg = threadIdx.y;
  const bool ok = c < C;
  const float frows = (float)rows;
  float a = 0.f;
  const float n_full = (float)rows_per_chunk, n_last = (float)(rows - (long long)(chunks - 1) * rows_per_chunk);
  if (ok) {
#pragma unroll 4
    for (int j = seg; j < chunks; j += kBnSegments)
      a = __fadd_rn(a, __fmul_rn(j == chunks - 1 ? n_last : n_full, __ldg(partial + ((long long)j * 2 + 0) * C + c)));
  }
  red[seg][threadIdx.x] = a;
  __syncthreads();
  if (seg == 0) {
    for (int g = 1; g < kBnSegments; ++g) a = __fadd_rn(a, red[g][threadIdx.x]);
    mean_s[threadIdx.x] = __fdiv_rn(a, frows);
  }
  __syncthreads();
  const float m = mean_s[threadIdx.x];
  float q = 0.f;
  if (ok) {
#pragma unroll 4
    for (int j = seg; j < chunks; j += kBnSegments) {
      const float d = __fsub_rn(__ldg(partial + ((long long)j * 2 + 0) * C + c), m);
      q = __fadd_rn(q, __fadd_rn(__ldg(partial + ((long long)j * 2 + 1) * C + c), __fmul_rn(j == chunks - 1 ? n_last : n_full, __fmul_rn(d, d))));
    }
  }
  __syncthreads();
  red[seg][threadIdx.x] = q;
  __syncthreads();
  if (seg != 0 || !ok) return;
  for (int g = 1; g < kBnSegments; ++g) q = __fadd_rn(q, red[g][threadIdx.x]);
  const float v = __fdiv_rn(q, frows);  // biased
  mean_out[c] = m;
  var_out[c] = v;
  if (update_moving) {
    float batch_var;
    if (fused && bn_mode == 0) batch_var = m;  // reference wiring: TakeOut index 1 twice (nn/AllKernels.scala:444-445)
    else if (fused) batch_var = __fmul_rn(v, __fdiv_rn(frows, __fsub_rn(frows, 1.f)));  // TF returns Bessel-corrected
    else batch_var = v;
    const float om = __fsub_rn(1.f, momentum);
    mov_mean[c] = __fadd_rn(__fmul_rn(momentum, mov_mean[c]), __fmul_rn(om, m));
    mov_var[c] = __fadd_rn(__fmul_rn(momentum, mov_var[c]), __fmul_rn(om, batch_var));
  }
}

struct BnChunks {
  int chunks;
  long long rows_per_chunk;
};
// chunks of whole block-rows x unroll, about three blocks per SM; bounded by the workspace ([chunk][2][C] floats)
BnChunks bn_chunks(long long rows, int C, size_t wsf, int unroll = kBnUnroll, int per_sm = 3) {
  const int R = kBnThreads / (C / 4);
  const long long unit = (long long)R * unroll;
  long long want = (long long)per_sm * kNumSMs;
  const long long by_ws = (long long)(wsf / ((size_t)2 * C));
  if (want > by_ws) want = by_ws;
  if (want < 1) want = 1;
  long long rpc = (rows + want - 1) / want;
  rpc = (rpc + unit - 1) / unit * unit;
  return BnChunks{(int)((rows + rpc - 1) / rpc), rpc};
}

// y = BN(x) with the per-channel constants of bn_apply4_kernel (kernels_simt.cu), four channels at a time
struct BnScale {
  float inv[4], mn[4], bt[4];
  int fused;
  __device__ __forceinline__ void load(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int c, int fused_) {
    const float4 vr = ldg4(var + c), gm = ldg4(gamma + c), m4 = ldg4(mean + c), b4 = ldg4(beta + c);
    const float v[4] = {vr.x, vr.y, vr.z, vr.w}, g[4] = {gm.x, gm.y, gm.z, gm.w};
    mn[0] = m4.x; mn[1] = m4.y; mn[2] = m4.z; mn[3] = m4.w;
    bt[0] = b4.x; bt[1] = b4.y; bt[2] = b4.z; bt[3] = b4.w;
#pragma unroll
    for (int k = 0; k < 4; ++k) inv[k] = __fmul_rn(__fdiv_rn(1.f, sqrtf(__fadd_rn(v[k], eps))), g[k]);
    fused = fused_;
  }
  __device__ __forceinline__ float apply(float xv, int k) const {
    return fused ? __fadd_rn(__fmul_rn(__fsub_rn(xv, mn[k]), inv[k]), bt[k])
                 : __fadd_rn(__fsub_rn(__fmul_rn(xv, inv[k]), __fmul_rn(mn[k], inv[k])), bt[k]);
  }
};

__device__ __forceinline__ unsigned first_max4(float a, float b, float c, float d, float* mx) {
  float best = a;
  unsigned idx = 0;
  if (b > best) { best = b; idx = 1; }
  if (c > best) { best = c; idx = 2; }
  if (d > best) { best = d; idx = 3; }
  *mx = best;
  return idx | 4u;  // bit 2 as in win4 without a ReLU behind the pool
}

// one thread per pooled (n, ph, pw, 4 channels)
__global__ void __launch_bounds__(kBnThreads) bn_apply_pool_kernel(const float* __restrict__ x, float* __restrict__ yp, unsigned char* __restrict__ map,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ mean, const float* __restrict__ var, float eps, int fused,
                                                                  int H, int W, int C4, int OH, int OW, long long n4) {
  pdl_prologue();
  // 32-bit index math (bn_pool_ok keeps every element count below 2^31: a 64-bit division costs more than the item's arithmetic)
  const unsigned stride = gridDim.x * kBnThreads;  // a multiple of C4: the channel group of a thread is fixed
  unsigned i = blockIdx.x * kBnThreads + threadIdx.x;
  const unsigned n = (unsigned)n4;
  if (i >= n) return;
  const int c4 = (int)(i % (unsigned)C4);
  BnScale bn;
  bn.load(gamma, beta, mean, var, eps, c4 * 4, fused);
  const float4* xp = reinterpret_cast<const float4*>(x) + c4;
  for (; i < n; i += stride) {
    unsigned p = i / (unsigned)C4;
    const unsigned t1 = p / (unsigned)OW, pw = p - t1 * (unsigned)OW;
    const unsigned img = t1 / (unsigned)OH, ph = t1 - img * (unsigned)OH;
    const unsigned base = ((img * H + 2 * ph) * W + 2 * pw) * C4;
    const float4 qa = __ldg(xp + base), qb = __ldg(xp + base + C4), qc = __ldg(xp + base + W * C4), qd = __ldg(xp + base + W * C4 + C4);
    float4 o;
    const unsigned b0 = first_max4(bn.apply(qa.x, 0), bn.apply(qb.x, 0), bn.apply(qc.x, 0), bn.apply(qd.x, 0), &o.x);
    const unsigned b1 = first_max4(bn.apply(qa.y, 1), bn.apply(qb.y, 1), bn.apply(qc.y, 1), bn.apply(qd.y, 1), &o.y);
    const unsigned b2 = first_max4(bn.apply(qa.z, 2), bn.apply(qb.z, 2), bn.apply(qc.z, 2), bn.apply(qd.z, 2), &o.z);
    const unsigned b3 = first_max4(bn.apply(qa.w, 3), bn.apply(qb.w, 3), bn.apply(qc.w, 3), bn.apply(qd.w, 3), &o.w);
    reinterpret_cast<float4*>(yp)[i] = o;
    if (map) reinterpret_cast<unsigned*>(map)[i] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
  }
}

// backward column sums over the POOLED positions: sum g and sum g * (x_winner - mean_in); partial[chunk][2][C]
__global__ void __launch_bounds__(kBnThreads) bn_pool_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dyp,
                                                                       const unsigned char* __restrict__ map, const float* __restrict__ save_mean,
                                                                       const float* __restrict__ save_var, float* __restrict__ partial, int H, int W,
                                                                       int C4, int OH, int OW, long long prow, long long rows_per_chunk, float frows,
                                                                       int quirk) {
  pdl_prologue();
  __shared__ float4 s0[kBnThreads];
  __shared__ float4 s1[kBnThreads];
  const int t = threadIdx.x, c4 = t % C4, j = t / C4, R = kBnThreads / C4;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = r0 + rows_per_chunk < prow ? r0 + rows_per_chunk : prow;
  float4 mu = ldg4(save_mean + c4 * 4);
  if (quirk) {  // reference wiring: the "mean" fed back is var_b * R/(R-1)
    const float4 v = ldg4(save_var + c4 * 4);
    const float k = __fdiv_rn(frows, __fsub_rn(frows, 1.f));
    mu = make_float4(__fmul_rn(v.x, k), __fmul_rn(v.y, k), __fmul_rn(v.z, k), __fmul_rn(v.w, k));
  }
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  const float4* xp = reinterpret_cast<const float4*>(x) + c4;
#pragma unroll 2
  for (unsigned r = (unsigned)r0 + j; r < (unsigned)r1; r += R) {
    const unsigned t1 = r / (unsigned)OW, pw = r - t1 * (unsigned)OW;
    const unsigned img = t1 / (unsigned)OH, ph = t1 - img * (unsigned)OH;
    const unsigned base = ((img * H + 2 * ph) * W + 2 * pw) * C4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(dyp) + r * C4 + c4);
    const unsigned mp = __ldg(reinterpret_cast<const unsigned*>(map) + r * C4 + c4);
    const float4 qa = __ldg(xp + base), qb = __ldg(xp + base + C4), qc = __ldg(xp + base + W * C4), qd = __ldg(xp + base + W * C4 + C4);
#define SB_PICK(f, sh) ((((mp >> (sh)) & 2u) ? (((mp >> (sh)) & 1u) ? qd.f : qc.f) : (((mp >> (sh)) & 1u) ? qb.f : qa.f)))
    const float xw0 = SB_PICK(x, 0), xw1 = SB_PICK(y, 8), xw2 = SB_PICK(z, 16), xw3 = SB_PICK(w, 24);
#undef SB_PICK
    a0.x = __fadd_rn(a0.x, g.x); a0.y = __fadd_rn(a0.y, g.y); a0.z = __fadd_rn(a0.z, g.z); a0.w = __fadd_rn(a0.w, g.w);
    a1.x = __fadd_rn(a1.x, __fmul_rn(g.x, __fsub_rn(xw0, mu.x))); a1.y = __fadd_rn(a1.y, __fmul_rn(g.y, __fsub_rn(xw1, mu.y)));
    a1.z = __fadd_rn(a1.z, __fmul_rn(g.z, __fsub_rn(xw2, mu.z))); a1.w = __fadd_rn(a1.w, __fmul_rn(g.w, __fsub_rn(xw3, mu.w)));
  }
  s0[t] = a0;
  s1[t] = a1;
  __syncthreads();
  if (j == 0) {
    for (int jj = 1; jj < R; ++jj) {
      const float4 u = s0[jj * C4 + c4], v = s1[jj * C4 + c4];
      a0.x = __fadd_rn(a0.x, u.x); a0.y = __fadd_rn(a0.y, u.y); a0.z = __fadd_rn(a0.z, u.z); a0.w = __fadd_rn(a0.w, u.w);
      a1.x = __fadd_rn(a1.x, v.x); a1.y = __fadd_rn(a1.y, v.y); a1.z = __fadd_rn(a1.z, v.z); a1.w = __fadd_rn(a1.w, v.w);
    }
    const int C = C4 * 4;
    *reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * 2 + 0) * C + c4 * 4) = a0;
    *reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * 2 + 1) * C + c4 * 4) = a1;
  }
}

// sums of the chunk partials (32 interleaved segments, then the segments in order) -> dbeta, dgamma and the two sums the apply kernel needs
__global__ void __launch_bounds__(32 * kBnSegments) bn_bwd_sums_kernel(const float* __restrict__ partial, int chunks, int C,
                                                                       const float* __restrict__ save_mean, const float* __restrict__ save_var,
                                                                       float eps, int quirk, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                       float* __restrict__ sums) {
  pdl_prologue();
  __shared__ float sa[kBnSegments][33], sc[kBnSegments][33];
  const int c = blockIdx.x * 32 + threadIdx.x, seg = threadIdx.y;
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
#pragma unroll 4
    for (int j = seg; j < chunks; j += kBnSegments) {
      a0 = __fadd_rn(a0, __ldg(partial + ((long long)j * 2 + 0) * C + c));
      a1 = __fadd_rn(a1, __ldg(partial + ((long long)j * 2 + 1) * C + c));
    }
  }
  sa[seg][threadIdx.x] = a0;
  sc[seg][threadIdx.x] = a1;
  __syncthreads();
  if (seg != 0 || c >= C) return;
  for (int g = 1; g < kBnSegments; ++g) {
    a0 = __fadd_rn(a0, sa[g][threadIdx.x]);
    a1 = __fadd_rn(a1, sc[g][threadIdx.x]);
  }
  const float var_in = quirk ? save_mean[c] : save_var[c];
  const float istd = __fdiv_rn(1.f, sqrtf(__fadd_rn(var_in, eps)));
  dbeta[c] = a0;
  dgamma[c] = __fmul_rn(a1, istd);
  sums[c] = a0;
  sums[C + c] = a1;
}

// dx for every input pixel: one thread per 2x2 block (n, bh, bw, 4 channels) of the ceil(H/2) x ceil(W/2) block grid; blocks past the
// pooled range (odd H or W) belong to no window, their dy is zero -- their dx is not.  relu: the in-place ReLU in front of the
// BatchNorm (x is its output): dx = [x > thr] * dx.
__global__ void __launch_bounds__(kBnThreads) bn_pool_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dyp,
                                                                      const unsigned char* __restrict__ map, float* __restrict__ dx,
                                                                      const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                                                      const float* __restrict__ save_var, const float* __restrict__ sums, float eps,
                                                                      float rows, int quirk, int relu, float thr, int H, int W, int C4, int OH, int OW,
                                                                      int BH, int BW, long long n4) {
  pdl_prologue();
  const unsigned stride = gridDim.x * kBnThreads;
  unsigned i = blockIdx.x * kBnThreads + threadIdx.x;
  const unsigned n = (unsigned)n4;
  if (i >= n) return;
  const int c4 = (int)(i % (unsigned)C4), c = c4 * 4, C = C4 * 4;
  float mean_in[4], ve[4], gi[4], mg[4], mgx[4];
  {
    const float4 m4 = ldg4(save_mean + c), v4 = ldg4(save_var + c), s0 = ldg4(sums + c), s1 = ldg4(sums + C + c), gm = ldg4(gamma + c);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w}, a[4] = {s0.x, s0.y, s0.z, s0.w},
                b[4] = {s1.x, s1.y, s1.z, s1.w}, g[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // the constants of bn_bwd_apply4_kernel (kernels_simt.cu)
      mean_in[k] = quirk ? __fmul_rn(v[k], __fdiv_rn(rows, __fsub_rn(rows, 1.f))) : m[k];
      ve[k] = __fadd_rn(quirk ? m[k] : v[k], eps);
      gi[k] = __fmul_rn(g[k], __fdiv_rn(1.f, sqrtf(ve[k])));
      mg[k] = __fdiv_rn(a[k], rows);
      mgx[k] = __fdiv_rn(__fdiv_rn(b[k], rows), ve[k]);  // mean(dy * xc) / (var_in + eps), folded once per channel
    }
  }
  const float4* xp = reinterpret_cast<const float4*>(x) + c4;
  float4* dxp = reinterpret_cast<float4*>(dx) + c4;
  for (; i < n; i += stride) {
    const unsigned p = i / (unsigned)C4;
    const unsigned t1 = p / (unsigned)BW, ubw = p - t1 * (unsigned)BW;
    const unsigned img = t1 / (unsigned)BH, ubh = t1 - img * (unsigned)BH;
    const int bw = (int)ubw, bh = (int)ubh;
    const bool inwin = bh < OH && bw < OW;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned mp = 0;
    if (inwin) {
      const unsigned pi = ((img * OH + bh) * OW + bw) * C4 + c4;
      const float4 q = __ldg(reinterpret_cast<const float4*>(dyp) + pi);
      g[0] = q.x; g[1] = q.y; g[2] = q.z; g[3] = q.w;
      mp = __ldg(reinterpret_cast<const unsigned*>(map) + pi) & 0x03030303u;
    } else {
      mp = 0x80808080u;  // matches no member
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      if (2 * bh + a >= H) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        if (2 * bw + b >= W) continue;
        const unsigned off = ((img * H + 2 * bh + a) * W + 2 * bw + b) * C4;
        const float4 q = __ldg(xp + off);
        const float xv[4] = {q.x, q.y, q.z, q.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float dy = ((mp >> (8 * k)) & 0xffu) == (unsigned)(a * 2 + b) ? g[k] : 0.f;
          const float xc = __fsub_rn(xv[k], mean_in[k]);
          const float tt = __fsub_rn(__fsub_rn(dy, mg[k]), __fmul_rn(xc, mgx[k]));
          const float r = __fmul_rn(gi[k], tt);
          o[k] = (!relu || xv[k] > thr) ? r : 0.f;
        }
        dxp[off] = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

int bn_grid(long long n4, int C4) {
  // grid * 256 must be a multiple of C4 (C4 divides 256, so any grid is); cap at 8 blocks per SM
  long long b = (n4 + kBnThreads - 1) / kBnThreads;
  const long long cap = (long long)kNumSMs * 8;
  (void)C4;
  return (int)std::max<long long>(1, std::min(b, cap));
}

}  // namespace

bool bn_lean_ok(int C) { return C % 4 == 0 && C / 4 <= kBnThreads && kBnThreads % (C / 4) == 0; }

bool bn_stats(const float* x, long long rows, int C, float* save_mean, float* save_var, float* mov_mean, float* mov_var, float momentum,
              int fused, int bn_mode, int update_moving, float* ws, size_t wsf, cudaStream_t s) {
  if (!bn_lean_ok(C) || ((uintptr_t)x % 16) != 0 || wsf < (size_t)2 * C) return false;
  const BnChunks cc = bn_chunks(rows, C, wsf);
  launch_k(bn_stats_kernel, cc.chunks, kBnThreads, 0, s, x, ws, rows, C / 4, cc.rows_per_chunk);
  SB_LAUNCHED();
  launch_k(bn_stats_finalize_kernel, ceil_div(C, 32), dim3(32, kBnSegments), 0, s, ws, cc.chunks, C, rows, cc.rows_per_chunk, save_mean, save_var,
           mov_mean, mov_var, momentum, fused, bn_mode, update_moving);
  SB_LAUNCHED();
  return true;
}

bool bn_pool_ok(const PoolGeom& g) {
  return bn_lean_ok(g.C) && g.KH == 2 && g.KW == 2 && g.SH == 2 && g.SW == 2 && g.PT == 0 && g.PL == 0 && g.OH == g.H / 2 && g.OW == g.W / 2 &&
         g.OH >= 1 && g.OW >= 1 && (long long)g.N * (g.H + 1) * (g.W + 1) * g.C < (1ll << 31);  // 32-bit element indices in the kernels
}

bool bn_apply_pool_fwd(const float* x, float* yp, unsigned char* map, const float* gamma, const float* beta, const float* mean, const float* var,
                       float eps, int fused, const PoolGeom& g, cudaStream_t s) {
  if (!bn_pool_ok(g) || (((uintptr_t)x | (uintptr_t)yp | (uintptr_t)map) % 16) != 0) return false;
  const int C4 = g.C / 4;
  const long long n4 = (long long)g.N * g.OH * g.OW * C4;
  launch_k(bn_apply_pool_kernel, bn_grid(n4, C4), kBnThreads, 0, s, x, yp, map, gamma, beta, mean, var, eps, fused, g.H, g.W, C4, g.OH, g.OW, n4);
  SB_LAUNCHED();
  return true;
}

bool bn_pool_bwd(const float* x, const float* dyp, const unsigned char* map, float* dx, const float* gamma, const float* save_mean,
                 const float* save_var, float* dgamma, float* dbeta, const PoolGeom& g, float eps, int fused, int bn_mode, int relu, float thr,
                 float* ws, size_t wsf, cudaStream_t s) {
  if (!bn_pool_ok(g) || !map || (((uintptr_t)x | (uintptr_t)dyp | (uintptr_t)map | (uintptr_t)dx) % 16) != 0 || wsf < (size_t)4 * g.C) return false;
  const int C = g.C, C4 = C / 4;
  const int quirk = (fused && bn_mode == 0) ? 1 : 0;
  const long long rows = (long long)g.N * g.H * g.W, prow = (long long)g.N * g.OH * g.OW;
  float* sums = ws;  // [2*C sums][partials...]
  float* part = ws + 2 * (size_t)C;
  BnChunks cc = bn_chunks(prow, C, wsf - 2 * (size_t)C, 1, 4);  // whole block-rows per chunk, four blocks per SM
  launch_k(bn_pool_bwd_reduce_kernel, cc.chunks, kBnThreads, 0, s, x, dyp, map, save_mean, save_var, part, g.H, g.W, C4, g.OH, g.OW, prow,
           cc.rows_per_chunk, (float)rows, quirk);
  SB_LAUNCHED();
  launch_k(bn_bwd_sums_kernel, ceil_div(C, 32), dim3(32, kBnSegments), 0, s, part, cc.chunks, C, save_mean, save_var, eps, quirk, dgamma, dbeta, sums);
  SB_LAUNCHED();
  if (dx) {
    const int BH = (g.H + 1) / 2, BW = (g.W + 1) / 2;
    const long long n4 = (long long)g.N * BH * BW * C4;
    launch_k(bn_pool_bwd_apply_kernel, bn_grid(n4, C4), kBnThreads, 0, s, x, dyp, map, dx, gamma, save_mean, save_var, sums, eps, (float)rows, quirk,
             relu, thr, g.H, g.W, C4, g.OH, g.OW, BH, BW, n4);
    SB_LAUNCHED();
  }
  return true;
}

void trace_set_bn_pool(unsigned long long* p) { sb_trace_set_local(p); }  // SB_TRACE timeline buffer of this translation unit

}  // namespace sb
