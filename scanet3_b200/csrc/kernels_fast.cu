// Vectorised HBM/L2-bound kernels for the shapes the BASELINE configs hit (128-bit accesses, channel-fastest NHWC):
//   * max-pool forward / backward (gather form, first-max rule) with the neighbouring ReLU folded into the backward,
//   * Conv2D with a tiny reduction (KH*KW*Cin <= 32, e.g. the first layer on 1- or 3-channel images): fprop (+ReLU) and wgrad,
//   * the skinny dense head (N <= 16 outputs): split-K forward, fused bias + softmax + CCE (+ bias gradient), wgrad, dgrad.
// All reductions use fixed orders (no atomics), so results are run-to-run deterministic.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_sm100.cuh"

namespace sb {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 max4(float4 a, float4 b) { return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)); }

// Winner byte of a 2x2 max-pool window (members a b / c d in row-major order): bits 0-1 = index of the FIRST maximum (a later
// member displaces an earlier one only when strictly greater -- TF's MaxPoolGrad rule, pinned by math/nn/KernelsSpec.scala:277-288),
// bit 2 = the maximum passes the preceding ReLU (max > thr; always set without a ReLU).  The backward pass routes dy with this
// byte alone: it never re-reads the activations.
__device__ __forceinline__ unsigned win4(float a, float b, float c, float d, int relu, float thr, float* mx) {
  float best = a;
  unsigned idx = 0;
  if (b > best) { best = b; idx = 1; }
  if (c > best) { best = c; idx = 2; }
  if (d > best) { best = d; idx = 3; }
  *mx = best;
  return idx | ((!relu || best > thr) ? 4u : 0u);
}
// 0xff in every byte lane of `m` whose winner byte routes to member `idx` (with the ReLU bit set when relu)
__device__ __forceinline__ unsigned take4(unsigned m, unsigned idx, int relu) {
  return relu ? __vcmpeq4(m, 0x01010101u * (idx | 4u)) : __vcmpeq4(m & 0x03030303u, 0x01010101u * idx);
}

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
  pdl_prologue();
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

// ---- 2x2 window, stride 1, VALID (the layer default, layer/Pool2D.scala:29-31): strip kernels ----
// A thread owns (image, output row, channel quad) and walks a segment of the row keeping the previous column's vertical
// maximum in registers: 2 loads per output instead of 4.
constexpr int kPoolSegs = 4;
__global__ void __launch_bounds__(256) pool22_fwd_strip_kernel(const float* __restrict__ x, float* __restrict__ y, PoolGeom g,
                                                               int seg_len, long long n_threads) {
  pdl_prologue_trigger();
  const int C4 = g.C >> 2;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;  // 32-bit index split (the launcher keeps n_threads < 2^31)
  if ((long long)i >= n_threads) return;
  const int c4 = (int)(i % (unsigned)C4);
  unsigned t = i / (unsigned)C4;
  const int seg = (int)(t % kPoolSegs); t /= kPoolSegs;
  const int oh = (int)(t % (unsigned)g.OH);
  const int b = (int)(t / (unsigned)g.OH);
  const int w0 = seg * seg_len;
  const int w1 = min(w0 + seg_len, g.OW);
  if (w0 >= w1) return;
  const float* r0 = x + (((long long)b * g.H + oh) * g.W) * g.C + c4 * 4;
  const float* r1 = r0 + (long long)g.W * g.C;
  float* yo = y + (((long long)b * g.OH + oh) * g.OW) * g.C + c4 * 4;
  r0 += w0 * g.C; r1 += w0 * g.C; yo += w0 * g.C;  // running pointers: one add per pointer and column
  float4 prev = max4(ld4(r0), ld4(r1));
#pragma unroll 4
  for (int w = w0; w < w1; ++w) {
    r0 += g.C; r1 += g.C;
    const float4 next = max4(ld4(r0), ld4(r1));
    st4(yo, max4(prev, next));
    yo += g.C;
    prev = next;
  }
}
// the same walk recording the winner byte of every window (win4) next to the maximum: the backward pass then routes dy by the map
// and never reads x again.  Keeps the previous column's two members in registers: still 2 loads per output.
__global__ void __launch_bounds__(256) pool22_fwd_map_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                             unsigned* __restrict__ map, PoolGeom g, int relu, float thr,
                                                             int seg_len, long long n_threads) {
  pdl_prologue_trigger();
  const int C4 = g.C >> 2;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if ((long long)i >= n_threads) return;
  const int c4 = (int)(i % (unsigned)C4);
  unsigned t = i / (unsigned)C4;
  const int seg = (int)(t % kPoolSegs); t /= kPoolSegs;
  const int oh = (int)(t % (unsigned)g.OH);
  const int b = (int)(t / (unsigned)g.OH);
  const int w0 = seg * seg_len;
  const int w1 = min(w0 + seg_len, g.OW);
  if (w0 >= w1) return;
  const float* r0 = x + (((long long)b * g.H + oh) * g.W) * g.C + c4 * 4;
  const float* r1 = r0 + (long long)g.W * g.C;
  float* yo = y + (((long long)b * g.OH + oh) * g.OW) * g.C + c4 * 4;
  unsigned* mo = map + (((long long)b * g.OH + oh) * g.OW) * C4 + c4;
  r0 += w0 * g.C; r1 += w0 * g.C; yo += w0 * g.C; mo += w0 * C4;
  float4 pa = ld4(r0), pc = ld4(r1);  // members 0 (top-left) and 2 (bottom-left) of the window anchored in column w
#pragma unroll 4
  for (int w = w0; w < w1; ++w) {
    r0 += g.C; r1 += g.C;
    const float4 nb = ld4(r0), nd = ld4(r1);  // members 1 and 3
    float4 m;
    const unsigned b0 = win4(pa.x, nb.x, pc.x, nd.x, relu, thr, &m.x);
    const unsigned b1 = win4(pa.y, nb.y, pc.y, nd.y, relu, thr, &m.y);
    const unsigned b2 = win4(pa.z, nb.z, pc.z, nd.z, relu, thr, &m.z);
    const unsigned b3 = win4(pa.w, nb.w, pc.w, nd.w, relu, thr, &m.w);
    st4(yo, m);
    *mo = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    yo += g.C; mo += C4;
    pa = nb; pc = nd;
  }
}
static bool pool_is_22s1(const PoolGeom& g) {
  return g.KH == 2 && g.KW == 2 && g.SH == 1 && g.SW == 1 && g.PT == 0 && g.PL == 0 && g.OH == g.H - 1 && g.OW == g.W - 1;
}
void pool_max_fwd_v4(const float* x, float* y, const PoolGeom& g, cudaStream_t s, unsigned char* map, int relu, float thr) {
  const long long nt22 = (long long)g.N * g.OH * kPoolSegs * (g.C / 4);  // 4 segments: 2, 3 and 6 measured 1.3-1.7 % slower on cfg2
  if (pool_is_22s1(g) && nt22 < (1LL << 31)) {  // the strip kernels split a 32-bit thread index
    const long long nt = nt22;
    const int seg_len = (g.OW + kPoolSegs - 1) / kPoolSegs;
    if (map) {
      launch_k(pool22_fwd_map_kernel, (int)((nt + 255) / 256), 256, 0, s, x, y, (unsigned*)map, g, relu, thr, seg_len, nt);
      SB_LAUNCHED();
      return;
    }
    launch_k(pool22_fwd_strip_kernel, (int)((nt + 255) / 256), 256, 0, s, x, y, g, seg_len, nt);
    SB_LAUNCHED();
    return;
  }
  const long long n4 = (long long)g.N * g.OH * g.OW * (g.C / 4);
  launch_k(pool_max_fwd_v4_kernel, grid_for(n4, 256), 256, 0, s, x, y, g, n4);
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
  pdl_prologue();
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

// Backward strip: a thread owns (image, INPUT row h, channel quad) and walks a segment of the row.  Column w of the walk
// loads column w+1 of the rows h-1, h, h+1 and the dy of the two windows whose top-left corner is in column w (3 x loads +
// 2 dy loads) and keeps columns w-1, w in registers; the element (h, w) takes dy of a window iff it is that window's FIRST
// maximum, decided by compares against maxima of its 8 neighbours.  Summation order per element: windows
// (h-1,w-1), (h-1,w), (h,w-1), (h,w) -- the same as the generic gather kernel.
//
// WG = KH*100 + KW*10 + CIN != 0 additionally fuses the FILTER GRADIENT of the small-K Conv2D that produced the pool's input
// (conv -> in-place ReLU -> pool, the first block of the BASELINE CNNs): the thread already holds dA = dLoss/d(conv output)
// for its pixel and 4 channels, so it accumulates dW[kh][kw][ci][co] += X[pixel + tap][ci] * dA[co] in registers over its strip
// (the conv input window slides with the walk: KH*CIN loads per column); threads of a CTA are then summed (shuffles + shared
// memory, fixed order) into one partial per CTA.  dA itself is only stored when something else needs it.
struct PoolWgradFuse {
  const float* xin;  // conv input [N, Hin, Win, CIN]
  float* partial;    // [gridDim.x][KH*KW*CIN][C]
  int Hin, Win, PT, PL;
  int store_dx;
};
// (the 3x3x1 fused form is capped at 128 registers: 4 CTAs per SM and 3 row segments on LeNet's 26-column rows; 40 bytes of spills)
template <bool RELU, int WG>
__global__ void __launch_bounds__(128, (WG == 331 ? 4 : 1)) pool22_bwd_strip_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dx, PoolGeom g, float thr, int seg_len,
                                                               int segs, long long n_threads, PoolWgradFuse wf) {
  pdl_prologue_trigger();
  constexpr int KH = WG / 100, KW = (WG / 10) % 10, CIN = WG % 10, K = KH * KW * CIN;
  const int C4 = g.C >> 2;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;  // 32-bit index split (the launcher keeps n_threads < 2^31)
  const int c4 = (int)(i % (unsigned)C4);
  unsigned t = i / (unsigned)C4;
  const int seg = (int)(t % (unsigned)segs); t /= (unsigned)segs;
  const int h = (int)(t % (unsigned)g.H);
  const int b = (int)(t / (unsigned)g.H);
  const int w0 = seg * seg_len;
  const int w1 = min(w0 + seg_len, g.W);
  const bool active = (long long)i < n_threads && w0 < w1;
  float4 dwacc[K > 0 ? K : 1];
  if (WG) {
#pragma unroll
    for (int k = 0; k < K; ++k) dwacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (active) {
    const bool top_ok = h >= 1, bot_ok = h <= g.H - 2;  // window rows h-1 and h exist
    const long long rs = (long long)g.W * g.C;
    const float* xm = x + (((long long)b * g.H + h) * g.W) * g.C + c4 * 4;  // row h
    const float* xt = top_ok ? xm - rs : xm;                                 // row h-1 (clamped: its windows then carry dy = 0)
    const float* xb = bot_ok ? xm + rs : xm;                                 // row h+1
    const float* dt = dy + (((long long)b * g.OH + (top_ok ? h - 1 : 0)) * g.OW) * g.C + c4 * 4;
    const float* db = dy + (((long long)b * g.OH + (bot_ok ? h : 0)) * g.OW) * g.C + c4 * 4;
    float* dxo = dx + (((long long)b * g.H + h) * g.W) * g.C + c4 * 4;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    // columns w-1 (p*), w (c*) of the rows h-1, h, h+1 and the dy of the two windows anchored in column w-1
    float4 pt, pm, pb, ct, cm, cb;
    float4 pgt = zero, pgb = zero;
    if (w0 >= 1) {
      const int o = (w0 - 1) * g.C;  // offsets inside one image row fit 32 bits
      pt = ld4(xt + o); pm = ld4(xm + o); pb = ld4(xb + o);
      ct = ld4(xt + o + g.C); cm = ld4(xm + o + g.C); cb = ld4(xb + o + g.C);
      pgt = top_ok ? ld4(dt + o) : zero;
      pgb = bot_ok ? ld4(db + o) : zero;
    } else {
      ct = ld4(xt); cm = ld4(xm); cb = ld4(xb);
      pt = ct; pm = cm; pb = cb;  // no windows in column -1: their dy is zero
    }
    // conv input window of the current pixel: rows h+a-PT, columns w+bb-PL (zero outside the image)
    float xw[KH > 0 ? KH : 1][KW > 0 ? KW : 1][CIN > 0 ? CIN : 1];
    auto load_col = [&](int col_in, int slot) {
#pragma unroll
      for (int a = 0; a < KH; ++a) {
        const int ih = h + a - wf.PT;
        const bool ok = ih >= 0 && ih < wf.Hin && col_in >= 0 && col_in < wf.Win;
        const float* p = wf.xin + (((long long)b * wf.Hin + (ok ? ih : 0)) * wf.Win + (ok ? col_in : 0)) * CIN;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) xw[a][slot][ci] = ok ? __ldg(p + ci) : 0.f;
      }
    };
    if (WG) {
#pragma unroll
      for (int bb = 0; bb + 1 < KW; ++bb) load_col(w0 - wf.PL + bb, bb + 1);  // shifted into place by the first iteration
    }
    // running pointers (one 64-bit add per pointer and column instead of an index multiply + sign extension each)
    const float *nxt = xt + (w0 + 1) * g.C, *nxm = xm + (w0 + 1) * g.C, *nxb = xb + (w0 + 1) * g.C;  // column w+1 of the three rows
    const float *dtw = dt + w0 * g.C, *dbw = db + w0 * g.C;                                          // dy of the windows of column w
    float* dxw = dxo + w0 * g.C;
#pragma unroll 2
    for (int w = w0; w < w1; ++w) {
      float4 gt = zero, gb = zero, nt = ct, nm = cm, nb = cb;
      if (w <= g.W - 2) {  // windows anchored in column w exist
        nt = ld4(nxt); nm = ld4(nxm); nb = ld4(nxb);
        gt = top_ok ? ld4(dtw) : zero;
        gb = bot_ok ? ld4(dbw) : zero;
      }
      nxt += g.C; nxm += g.C; nxb += g.C; dtw += g.C; dbw += g.C;
      // the element is the FIRST maximum of a window (row-major order, strict > displaces an earlier member) iff it beats the
      // earlier members strictly and the later ones weakly: compares against maxima of its 8 neighbours (FMNMX / FMNMX3) instead
      // of two 4-way arg-maxes with index bookkeeping
      float4 acc = zero;
#define SB_WIN(c)                                                                                                        \
  {                                                                                                                      \
    const float e = cm.c;                                                                                                \
    const float m1 = fmaxf(fmaxf(pt.c, ct.c), pm.c), m2 = fmaxf(ct.c, nt.c);                                             \
    const float m3 = fmaxf(pb.c, cb.c), m4 = fmaxf(fmaxf(nm.c, cb.c), nb.c);                                             \
    if (e > m1) acc.c = __fadd_rn(acc.c, pgt.c);                /* window (h-1, w-1): the bottom-right member */          \
    if (e > m2 && e >= nm.c) acc.c = __fadd_rn(acc.c, gt.c);    /* window (h-1, w)  : bottom-left             */          \
    if (e > pm.c && e >= m3) acc.c = __fadd_rn(acc.c, pgb.c);   /* window (h,   w-1): top-right               */          \
    if (e >= m4) acc.c = __fadd_rn(acc.c, gb.c);                /* window (h,   w)  : top-left                */          \
  }
      SB_WIN(x) SB_WIN(y) SB_WIN(z) SB_WIN(w)
#undef SB_WIN
      if (RELU) {
        if (!(cm.x > thr)) acc.x = 0.f;
        if (!(cm.y > thr)) acc.y = 0.f;
        if (!(cm.z > thr)) acc.z = 0.f;
        if (!(cm.w > thr)) acc.w = 0.f;
      }
      if (!WG || wf.store_dx) st4(dxw, acc);
      dxw += g.C;
      if (WG) {
#pragma unroll
        for (int a = 0; a < KH; ++a)
#pragma unroll
          for (int bb = 0; bb + 1 < KW; ++bb)
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) xw[a][bb][ci] = xw[a][bb + 1][ci];
        load_col(w - wf.PL + KW - 1, KW - 1);
#pragma unroll
        for (int a = 0; a < KH; ++a)
#pragma unroll
          for (int bb = 0; bb < KW; ++bb)
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
              const float xv = xw[a][bb][ci];
              float4& r = dwacc[(a * KW + bb) * CIN + ci];
              r.x = fmaf(xv, acc.x, r.x); r.y = fmaf(xv, acc.y, r.y); r.z = fmaf(xv, acc.z, r.z); r.w = fmaf(xv, acc.w, r.w);
            }
      }
      pgt = gt; pgb = gb;
      pt = ct; pm = cm; pb = cb;
      ct = nt; cm = nm; cb = nb;
    }
  }
  if (WG) {
    // CTA-level sum in a fixed order: lanes with the same channel quad inside a warp (xor C4, 2*C4, ..), then the warps
    extern __shared__ __align__(16) float4 red[];  // [nwarps][K][C4]
    const int wl = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cq = threadIdx.x % C4;  // == c4 when C4 divides the CTA size (checked by the launcher)
#pragma unroll
    for (int k = 0; k < K; ++k)
      for (int o = C4; o < 32; o <<= 1) {
        dwacc[k].x = __fadd_rn(dwacc[k].x, __shfl_xor_sync(0xffffffffu, dwacc[k].x, o));
        dwacc[k].y = __fadd_rn(dwacc[k].y, __shfl_xor_sync(0xffffffffu, dwacc[k].y, o));
        dwacc[k].z = __fadd_rn(dwacc[k].z, __shfl_xor_sync(0xffffffffu, dwacc[k].z, o));
        dwacc[k].w = __fadd_rn(dwacc[k].w, __shfl_xor_sync(0xffffffffu, dwacc[k].w, o));
      }
    if (wl < C4) {
#pragma unroll
      for (int k = 0; k < K; ++k) red[(wid * K + k) * C4 + cq] = dwacc[k];
    }
    __syncthreads();
    float* part = wf.partial + (size_t)blockIdx.x * K * g.C;
    for (int q = threadIdx.x; q < K * C4; q += blockDim.x) {
      float4 tot = red[q];
      for (int wq = 1; wq < nwarps; ++wq) {
        const float4 v = red[wq * K * C4 + q];
        tot.x = __fadd_rn(tot.x, v.x); tot.y = __fadd_rn(tot.y, v.y); tot.z = __fadd_rn(tot.z, v.z); tot.w = __fadd_rn(tot.w, v.w);
      }
      st4(part + (q / C4) * g.C + (q % C4) * 4, tot);
    }
  }
}
// The same walk driven by the forward pass's winner bytes (win4) instead of the activations: column w loads the dy and the map
// word of the two windows anchored in column w (4 loads, no activation rows) and keeps column w-1's in registers; element (h, w)
// takes the dy of window (h-1,w-1) iff that window's winner is its member 3 (bottom-right), of (h-1,w) iff member 2, of (h,w-1)
// iff member 1, of (h,w) iff member 0 -- and, under a ReLU, only when the winner passed it (bit 2).  Same summation order as
// the activation-driven kernels, hence the same bits.
template <int WG>
__global__ void __launch_bounds__(128, (WG == 331 ? 4 : 1)) pool22_bwd_map_kernel(const unsigned* __restrict__ map, const float* __restrict__ dy,
                                                               float* __restrict__ dx, PoolGeom g, int relu, int seg_len,
                                                               int segs, long long n_threads, PoolWgradFuse wf) {
  pdl_prologue_trigger();
  constexpr int KH = WG / 100, KW = (WG / 10) % 10, CIN = WG % 10, K = KH * KW * CIN;
  const int C4 = g.C >> 2;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;  // 32-bit index split (the launcher keeps n_threads < 2^31)
  const int c4 = (int)(i % (unsigned)C4);
  unsigned t = i / (unsigned)C4;
  const int seg = (int)(t % (unsigned)segs); t /= (unsigned)segs;
  const int h = (int)(t % (unsigned)g.H);
  const int b = (int)(t / (unsigned)g.H);
  const int w0 = seg * seg_len;
  const int w1 = min(w0 + seg_len, g.W);
  const bool active = (long long)i < n_threads && w0 < w1;
  float4 dwacc[K > 0 ? K : 1];
  if (WG) {
#pragma unroll
    for (int k = 0; k < K; ++k) dwacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (active) {
    const bool top_ok = h >= 1, bot_ok = h <= g.H - 2;  // window rows h-1 and h exist
    const float* dt = dy + (((long long)b * g.OH + (top_ok ? h - 1 : 0)) * g.OW) * g.C + c4 * 4;
    const float* db = dy + (((long long)b * g.OH + (bot_ok ? h : 0)) * g.OW) * g.C + c4 * 4;
    const unsigned* mt = map + (((long long)b * g.OH + (top_ok ? h - 1 : 0)) * g.OW) * C4 + c4;
    const unsigned* mb = map + (((long long)b * g.OH + (bot_ok ? h : 0)) * g.OW) * C4 + c4;
    float* dxo = dx + (((long long)b * g.H + h) * g.W) * g.C + c4 * 4;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    // windows anchored in column w-1: dy and winner words of the rows h-1 (t) and h (b); 0xffffffff routes nothing
    float4 pgt = zero, pgb = zero;
    unsigned pmt = 0xffffffffu, pmb = 0xffffffffu;
    if (w0 >= 1) {
      const int o = (w0 - 1) * g.C, om = (w0 - 1) * C4;
      if (top_ok) { pgt = ld4(dt + o); pmt = __ldg(mt + om); }
      if (bot_ok) { pgb = ld4(db + o); pmb = __ldg(mb + om); }
    }
    float xw[KH > 0 ? KH : 1][KW > 0 ? KW : 1][CIN > 0 ? CIN : 1];
    auto load_col = [&](int col_in, int slot) {
#pragma unroll
      for (int a = 0; a < KH; ++a) {
        const int ih = h + a - wf.PT;
        const bool ok = ih >= 0 && ih < wf.Hin && col_in >= 0 && col_in < wf.Win;
        const float* p = wf.xin + (((long long)b * wf.Hin + (ok ? ih : 0)) * wf.Win + (ok ? col_in : 0)) * CIN;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) xw[a][slot][ci] = ok ? __ldg(p + ci) : 0.f;
      }
    };
    if (WG) {
#pragma unroll
      for (int bb = 0; bb + 1 < KW; ++bb) load_col(w0 - wf.PL + bb, bb + 1);  // shifted into place by the first iteration
    }
    const float *dtw = dt + w0 * g.C, *dbw = db + w0 * g.C;
    const unsigned *mtw = mt + w0 * C4, *mbw = mb + w0 * C4;
    float* dxw = dxo + w0 * g.C;
#pragma unroll 2
    for (int w = w0; w < w1; ++w) {
      float4 gt = zero, gb = zero;
      unsigned wt = 0xffffffffu, wb = 0xffffffffu;
      if (w <= g.W - 2) {  // windows anchored in column w exist
        if (top_ok) { gt = ld4(dtw); wt = __ldg(mtw); }
        if (bot_ok) { gb = ld4(dbw); wb = __ldg(mbw); }
      }
      dtw += g.C; dbw += g.C; mtw += C4; mbw += C4;
      const unsigned e3 = take4(pmt, 3u, relu), e2 = take4(wt, 2u, relu), e1 = take4(pmb, 1u, relu), e0 = take4(wb, 0u, relu);
      float4 acc = zero;
#define SB_WIN(c, sh)                                            \
  if (e3 & (0xffu << sh)) acc.c = __fadd_rn(acc.c, pgt.c);       \
  if (e2 & (0xffu << sh)) acc.c = __fadd_rn(acc.c, gt.c);        \
  if (e1 & (0xffu << sh)) acc.c = __fadd_rn(acc.c, pgb.c);       \
  if (e0 & (0xffu << sh)) acc.c = __fadd_rn(acc.c, gb.c);
      SB_WIN(x, 0) SB_WIN(y, 8) SB_WIN(z, 16) SB_WIN(w, 24)
#undef SB_WIN
      if (!WG || wf.store_dx) st4(dxw, acc);
      dxw += g.C;
      if (WG) {
#pragma unroll
        for (int a = 0; a < KH; ++a)
#pragma unroll
          for (int bb = 0; bb + 1 < KW; ++bb)
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) xw[a][bb][ci] = xw[a][bb + 1][ci];
        load_col(w - wf.PL + KW - 1, KW - 1);
#pragma unroll
        for (int a = 0; a < KH; ++a)
#pragma unroll
          for (int bb = 0; bb < KW; ++bb)
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
              const float xv = xw[a][bb][ci];
              float4& r = dwacc[(a * KW + bb) * CIN + ci];
              r.x = fmaf(xv, acc.x, r.x); r.y = fmaf(xv, acc.y, r.y); r.z = fmaf(xv, acc.z, r.z); r.w = fmaf(xv, acc.w, r.w);
            }
      }
      pgt = gt; pgb = gb; pmt = wt; pmb = wb;
    }
  }
  if (WG) {
    // CTA-level sum in a fixed order: lanes with the same channel quad inside a warp (xor C4, 2*C4, ..), then the warps
    extern __shared__ __align__(16) float4 red[];  // [nwarps][K][C4]
    const int wl = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cq = threadIdx.x % C4;  // == c4 when C4 divides the CTA size (checked by the launcher)
#pragma unroll
    for (int k = 0; k < K; ++k)
      for (int o = C4; o < 32; o <<= 1) {
        dwacc[k].x = __fadd_rn(dwacc[k].x, __shfl_xor_sync(0xffffffffu, dwacc[k].x, o));
        dwacc[k].y = __fadd_rn(dwacc[k].y, __shfl_xor_sync(0xffffffffu, dwacc[k].y, o));
        dwacc[k].z = __fadd_rn(dwacc[k].z, __shfl_xor_sync(0xffffffffu, dwacc[k].z, o));
        dwacc[k].w = __fadd_rn(dwacc[k].w, __shfl_xor_sync(0xffffffffu, dwacc[k].w, o));
      }
    if (wl < C4) {
#pragma unroll
      for (int k = 0; k < K; ++k) red[(wid * K + k) * C4 + cq] = dwacc[k];
    }
    __syncthreads();
    float* part = wf.partial + (size_t)blockIdx.x * K * g.C;
    for (int q = threadIdx.x; q < K * C4; q += blockDim.x) {
      float4 tot = red[q];
      for (int wq = 1; wq < nwarps; ++wq) {
        const float4 v = red[wq * K * C4 + q];
        tot.x = __fadd_rn(tot.x, v.x); tot.y = __fadd_rn(tot.y, v.y); tot.z = __fadd_rn(tot.z, v.z); tot.w = __fadd_rn(tot.w, v.w);
      }
      st4(part + (q / C4) * g.C + (q % C4) * 4, tot);
    }
  }
}
// =====================================================================================================================
// Pool2D(2x2, stride 1) backward fused with the INPUT GRADIENT of the classifier head that consumes the pooled tensor
// (Pool2D >> Flatten >> Dense(N <= 16), layer/Dense.scala:57-61): dP[b][f] = sum_n G[b][n] W[f][n] is formed on the fly and routed
// through the winner map, so the head's dX (one write + one read of the whole pooled tensor) never exists and the head's backward
// kernel shrinks to its weight gradient (off the critical path, on the side branch).
//   CTA = (output row h of the pool's input, group of images).  Thread (r, pw, c4) owns the pooled position (h-1+r, pw) and four
//   channels: its 4 x N slab of W stays in REGISTERS for all the images of the group (W-stationary); per image it forms the dP
//   quad (the FFMA2 pairing and summation order of dense_head_bwd_kernel, so the bits are the same as the unfused path's) into a
//   double-buffered shared tile [2 rows][OW][C]; after one barrier the threads gather row h: element (h, w) takes dP of windows
//   (h-1,w-1), (h-1,w), (h,w-1), (h,w) iff their winner byte says so (same order as pool22_bwd_map_kernel).  The winner words of
//   image b+1 are fetched while image b computes.
// =====================================================================================================================
constexpr int kPoolHeadThreads = 704;  // 22 warps x 80 registers (warp allocations come in 512-register units: 88 would not launch)
// NC: compile-time class count (the thread's 4*NC weights are read with 128-bit shared loads), 0 = run-time N (scalar loads)
template <int NT, int NC>
__global__ void __launch_bounds__(kPoolHeadThreads, 1) pool22_bwd_head_kernel(const unsigned* __restrict__ map, const float* __restrict__ gmat,
                                                                            const float* __restrict__ w, float* __restrict__ dx,
                                                                            PoolGeom g, int N, int relu, int img_per_cta, int prewait) {
  extern __shared__ __align__(16) float psm[];
  const int C4 = g.C >> 2, npos = g.OW * C4;
  float* sg = psm;                                      // [img_per_cta][NT]
  float* dp = psm + (size_t)img_per_cta * NT;           // [2 buffers][2 rows][OW][C]
  const int h = blockIdx.x;
  const int b0 = blockIdx.y * img_per_cta, b1 = min(g.N, b0 + img_per_cta);
  const int tid = threadIdx.x;
  const int r = tid / npos, pos = tid - r * npos;       // r: 0 = pooled row h-1, 1 = pooled row h (tid >= 2*npos: gather only)
  const int pw = pos / C4, c4 = pos - pw * C4;
  const int ph = h - 1 + r;
  const bool owns = r < 2 && ph >= 0 && ph < g.OH;
  // W slab of the position: rows f = (ph*OW + pw)*C + c4*4 + q, q = 0..3, N classes each (zero beyond N) = 4*N contiguous floats.
  // The CTA's positions cover one contiguous piece of W (pooled rows h-1 and h): it is copied to shared memory with coalesced
  // 128-bit loads first -- per-thread gathers straight from global memory touch 32 different lines per warp instruction and cost
  // the LSU ~17 us per CTA (measured: 33 us for the whole kernel).
  if (!prewait) pdl_wait();
  float2 wr2[4][NT / 2];
  {
    const int ph_lo = max(h - 1, 0), ph_hi = min(h, g.OH - 1);
    const long long slab0 = (long long)ph_lo * npos * 4 * N;                    // floats; multiple of 4 (C % 4 == 0)
    const int slab_n = (ph_hi - ph_lo + 1) * npos * 4 * N;
    float* wsm = dp + (size_t)2 * 2 * g.OW * g.C;                                // [<= 2 rows][OW][C][N]
    const float4* src = reinterpret_cast<const float4*>(w + slab0);
    const int n4 = slab_n / 4;
    for (int i0 = tid; i0 < n4; i0 += 8 * blockDim.x) {  // 8 independent loads in flight per thread (a load->store chain per
      float4 t[8];                                        // element would pay the L2 latency ten times over)
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = i0 + u * blockDim.x < n4 ? __ldg(src + i0 + u * blockDim.x) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * blockDim.x < n4) reinterpret_cast<float4*>(wsm)[i0 + u * blockDim.x] = t[u];
    }
    __syncthreads();
    const float* wp = wsm + (size_t)((owns ? ph - ph_lo : 0) * npos + pos) * 4 * N;
    if (NC > 0) {  // 4*NC contiguous floats, 16-byte aligned: NC 128-bit loads (scalar loads at this 160-byte lane stride are 8-way bank conflicts)
      float wl[4 * (NC > 0 ? NC : 1)];
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(wp + 4 * i);
        wl[4 * i] = t.x; wl[4 * i + 1] = t.y; wl[4 * i + 2] = t.z; wl[4 * i + 3] = t.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int n = 0; n < NT / 2; ++n)
          wr2[q][n] = make_float2((owns && 2 * n < NC) ? wl[q * NC + 2 * n] : 0.f, (owns && 2 * n + 1 < NC) ? wl[q * NC + 2 * n + 1] : 0.f);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int n = 0; n < NT / 2; ++n)
          wr2[q][n] = make_float2((owns && 2 * n < N) ? wp[q * N + 2 * n] : 0.f, (owns && 2 * n + 1 < N) ? wp[q * N + 2 * n + 1] : 0.f);
    }
  }
  if (prewait) pdl_wait();
  pdl_trigger();
  for (int i = tid; i < (b1 - b0) * NT; i += blockDim.x) {
    const int m = i / NT, n = i - m * NT;
    sg[i] = n < N ? __ldg(gmat + (size_t)(b0 + m) * N + n) : 0.f;
  }
  // gather item of this thread: output pixel (h, gw), channel quad gc (one item per thread; W*C4 <= blockDim, checked by the launcher)
  const int gw = tid / C4, gc = tid - gw * C4;
  const bool gathers = gw < g.W;
  const bool top_ok = h >= 1, bot_ok = h <= g.H - 2, lft_ok = gw >= 1, rgt_ok = gw <= g.W - 2;
  // everything that does not depend on the image is hoisted: the four winner words' pointers (advanced by one image per
  // iteration; a window that does not exist reads a valid dummy address and is masked), the dP tile slots, the output pointer.
  // (The first version recomputed the 64-bit indices inside the loop: 78 of its ~250 instructions per image, ncu source page.)
  const bool v3 = gathers && top_ok && lft_ok, v2 = gathers && top_ok && rgt_ok, v1 = gathers && bot_ok && lft_ok, v0 = gathers && bot_ok && rgt_ok;
  // The winner words of the CTA's images (pooled rows h-1 and h: one contiguous block of <= 2*OW*C4 words per image) are copied
  // to shared memory up front with independent coalesced loads, so that the per-image loop below touches global memory only to
  // store dx (a one-image-ahead register prefetch left every iteration waiting for an L2 round trip).
  const int ph_lo = max(h - 1, 0), ph_hi = min(h, g.OH - 1);
  const int blk_words = (ph_hi - ph_lo + 1) * npos;         // words per image block
  const int map_img = g.OH * g.OW * C4;                     // words per image (the launcher keeps the whole map below 2^31 words)
  unsigned* smap = reinterpret_cast<unsigned*>(dp + (size_t)2 * 2 * g.OW * g.C + (size_t)2 * g.OW * g.C * N);  // [img_per_cta][2*npos]
  {
    const unsigned* msrc = map + (long long)b0 * map_img + ph_lo * npos;
    const int nimg = b1 - b0;
    if (tid < blk_words) {
#pragma unroll 4
      for (int m = 0; m < nimg; ++m) smap[m * 2 * npos + tid] = __ldg(msrc + (long long)m * map_img + tid);
    }
  }
  // word of window (row r in {h-1, h}, column c) of image m: smap[m*2*npos + (r - ph_lo)*npos + c*C4 + gc]
  const int o3 = v3 ? ((h - 1 - ph_lo) * g.OW + gw - 1) * C4 + gc : 0;  // window (h-1, w-1): member 3
  const int o2 = v2 ? ((h - 1 - ph_lo) * g.OW + gw) * C4 + gc : 0;      // window (h-1, w)  : member 2
  const int o1 = v1 ? ((h - ph_lo) * g.OW + gw - 1) * C4 + gc : 0;      // window (h,   w-1): member 1
  const int o0 = v0 ? ((h - ph_lo) * g.OW + gw) * C4 + gc : 0;          // window (h,   w)  : member 0
  const unsigned none = 0xffffffffu;
  const int tile_floats = 2 * g.OW * g.C;
  float* my_slot = dp + ((size_t)(r < 2 ? r : 0) * g.OW + pw) * g.C + c4 * 4;  // + buf * tile_floats
  const float* t3 = dp + (v3 ? (gw - 1) * g.C : 0) + gc * 4;                   // pooled row h-1
  const float* t2 = dp + (v2 ? gw * g.C : 0) + gc * 4;
  const float* t1 = dp + (size_t)g.OW * g.C + (v1 ? (gw - 1) * g.C : 0) + gc * 4;  // pooled row h
  const float* t0 = dp + (size_t)g.OW * g.C + (v0 ? gw * g.C : 0) + gc * 4;
  float* dxo = dx + (((long long)b0 * g.H + h) * g.W + (gathers ? gw : 0)) * g.C + gc * 4;
  const long long dx_img = (long long)g.H * g.W * g.C;
  const float* gr = sg;
  const unsigned* sm_words = smap;
  __syncthreads();
  int toff = 0;
  for (int b = b0; b < b1; ++b) {
    if (r < 2) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (owns) {
        float2 d2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) d2[q] = make_float2(0.f, 0.f);
#pragma unroll
        for (int n4 = 0; n4 < NT; n4 += 4) {
          const float4 gv = *reinterpret_cast<const float4*>(gr + n4);
          const float2 g0 = make_float2(gv.x, gv.y), g1 = make_float2(gv.z, gv.w);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            d2[q] = __ffma2_rn(g0, wr2[q][n4 / 2], d2[q]);
            d2[q] = __ffma2_rn(g1, wr2[q][n4 / 2 + 1], d2[q]);
          }
        }
        o = make_float4(__fadd_rn(d2[0].x, d2[0].y), __fadd_rn(d2[1].x, d2[1].y), __fadd_rn(d2[2].x, d2[2].y), __fadd_rn(d2[3].x, d2[3].y));
      }
      *reinterpret_cast<float4*>(my_slot + toff) = o;
    }
    gr += NT;
    __syncthreads();
    if (gathers) {
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 pgt = v3 ? *reinterpret_cast<const float4*>(t3 + toff) : zero;
      const float4 gt = v2 ? *reinterpret_cast<const float4*>(t2 + toff) : zero;
      const float4 pgb = v1 ? *reinterpret_cast<const float4*>(t1 + toff) : zero;
      const float4 gb = v0 ? *reinterpret_cast<const float4*>(t0 + toff) : zero;
      const unsigned m3 = v3 ? sm_words[o3] : none, m2 = v2 ? sm_words[o2] : none, m1 = v1 ? sm_words[o1] : none,
                     m0 = v0 ? sm_words[o0] : none;
      const unsigned e3 = take4(m3, 3u, relu), e2 = take4(m2, 2u, relu), e1 = take4(m1, 1u, relu), e0 = take4(m0, 0u, relu);
      float4 acc = zero;
#define SB_WIN(c, sh)                                            \
  if (e3 & (0xffu << sh)) acc.c = __fadd_rn(acc.c, pgt.c);       \
  if (e2 & (0xffu << sh)) acc.c = __fadd_rn(acc.c, gt.c);        \
  if (e1 & (0xffu << sh)) acc.c = __fadd_rn(acc.c, pgb.c);       \
  if (e0 & (0xffu << sh)) acc.c = __fadd_rn(acc.c, gb.c);
      SB_WIN(x, 0) SB_WIN(y, 8) SB_WIN(z, 16) SB_WIN(w, 24)
#undef SB_WIN
      st4(dxo, acc);
      dxo += dx_img;
    }
    toff = tile_floats - toff;  // the other buffer
    sm_words += 2 * npos;
  }
}
// false: shape not covered (the caller runs the head's dgrad and the plain pool backward)
bool pool_bwd_head_dgrad(const unsigned char* map, const float* gmat, const float* w, float* dx, const PoolGeom& g, int N, int relu,
                         cudaStream_t s) {
  if (!map || !pool_is_22s1(g) || (g.C & 3) || N > 16 || N < 1) return false;
  const int C4 = g.C / 4;
  if ((long long)g.N * g.OH * g.OW * C4 >= (1LL << 31)) return false;
  if (2 * g.OW * C4 > kPoolHeadThreads || g.W * C4 > kPoolHeadThreads) return false;
  const int NT = N <= 12 ? 12 : 16;
  // one wave of CTAs: rows x image groups ~ 148
  int groups = std::max(1, std::min(g.N, kNumSMs / std::max(1, g.H)));
  const int img_per_cta = (g.N + groups - 1) / groups;
  groups = (g.N + img_per_cta - 1) / img_per_cta;
  const size_t smem = ((size_t)img_per_cta * NT + (size_t)2 * 2 * g.OW * g.C + (size_t)2 * g.OW * g.C * N +
                       (size_t)img_per_cta * 2 * g.OW * C4) * sizeof(float);
  if (smem > 200 * 1024 || (((size_t)img_per_cta * NT) & 3)) return false;
  static bool attr = false;
  if (!attr) {
    SB_CUDA(cudaFuncSetAttribute(pool22_bwd_head_kernel<12, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SB_CUDA(cudaFuncSetAttribute(pool22_bwd_head_kernel<12, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SB_CUDA(cudaFuncSetAttribute(pool22_bwd_head_kernel<16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const int thr = std::max(2 * g.OW * C4, g.W * C4);
  const int threads = (thr + 31) / 32 * 32;
  const dim3 grid(g.H, groups);
  const int pw = prewait_ok();
  if (N == 10) launch_k(pool22_bwd_head_kernel<12, 10>, grid, threads, smem, s, (const unsigned*)map, gmat, w, dx, g, N, relu, img_per_cta, pw);
  else if (NT == 12) launch_k(pool22_bwd_head_kernel<12, 0>, grid, threads, smem, s, (const unsigned*)map, gmat, w, dx, g, N, relu, img_per_cta, pw);
  else launch_k(pool22_bwd_head_kernel<16, 0>, grid, threads, smem, s, (const unsigned*)map, gmat, w, dx, g, N, relu, img_per_cta, pw);
  SB_LAUNCHED();
  return true;
}

// segments per row: cost ~ whole waves x (columns per thread + prologue columns); a wave = what 148 SMs hold at the kernel's
// register footprint.  The prologue (index split, pointer setup, the column before the segment) costs ~2.5 columns of the walk
// (SASS: 315 instructions against 119 per column): LeNet's pool2 backward (23 columns) then takes 2 segments in one wave instead
// of 5 in two (933.5 k -> 937.7 k samples/s)
static void pool22_segments(const PoolGeom& g, int ctas_per_sm, int* segs_out, int* seg_len_out) {
  const long long wave = (long long)kNumSMs * ctas_per_sm * 128;
  int best_segs = 1;
  double best = 1e30;
  for (int segs = 1; segs <= 8; ++segs) {
    const int len = (g.W + segs - 1) / segs;
    if ((segs - 1) * len >= g.W) continue;
    const long long nt = (long long)g.N * g.H * segs * (g.C / 4);
    const double cost = (double)((nt + wave - 1) / wave) * (len + 2.5);
    if (cost < best) { best = cost; best_segs = segs; }
  }
  *segs_out = best_segs;
  *seg_len_out = (g.W + best_segs - 1) / best_segs;
}
static bool pool_is_22s1(const PoolGeom& g);
// max-pool backward (+ReLU mask) with the filter gradient of the preceding small-K conv: returns false when not covered
bool pool_bwd_conv_wgrad(const float* x, const float* dy, float* dx, const PoolGeom& g, int relu, float thr, const float* conv_in,
                         const ConvGeom& cg, float* dw, int store_dx, float* ws, size_t ws_floats, cudaStream_t s,
                         const unsigned char* map, ReduceBatch* defer) {
  const int C4 = g.C / 4, shape = cg.KH * 100 + cg.KW * 10 + cg.Cin;
  if (!pool_is_22s1(g) || (g.C & 3) || C4 > 32 || (C4 & (C4 - 1)) || cg.SH != 1 || cg.SW != 1) return false;
  if (cg.OH != g.H || cg.OW != g.W || cg.Cout != g.C || cg.N != g.N) return false;
  if (shape != 331 && shape != 333 && shape != 551) return false;
  int segs, seg_len;
  pool22_segments(g, shape == 331 ? 4 : 3, &segs, &seg_len);  // resident CTAs per SM of the instantiation
  const long long nt = (long long)g.N * g.H * segs * C4;
  if (nt >= (1LL << 31)) return false;  // the strip kernel splits a 32-bit thread index
  const int blocks = (int)((nt + 127) / 128);
  const int K = cg.KH * cg.KW * cg.Cin;
  if ((size_t)blocks * K * g.C > ws_floats) return false;
  const size_t smem = (size_t)4 * K * C4 * sizeof(float4);
  PoolWgradFuse wf{conv_in, ws, cg.H, cg.W, cg.PT, cg.PL, store_dx};
  if (map) {  // routed by the forward pass's winner bytes: the activations are not read (and were never written)
#define SB_POOL_MAP(S) \
  launch_k(pool22_bwd_map_kernel<S>, blocks, 128, smem, s, (const unsigned*)map, dy, dx, g, relu, seg_len, segs, nt, wf)
    if (shape == 331) SB_POOL_MAP(331); else if (shape == 333) SB_POOL_MAP(333); else SB_POOL_MAP(551);
#undef SB_POOL_MAP
    SB_LAUNCHED();
    if (defer && defer->add(ws, dw, K * g.C, blocks)) return true;  // summed with the step's other partials in one launch
    partial_reduce(ws, dw, K * g.C, blocks, s);
    return true;
  }
#define SB_POOL_WG(R, S) \
  launch_k(pool22_bwd_strip_kernel<R, S>, blocks, 128, smem, s, x, dy, dx, g, thr, seg_len, segs, nt, wf)
  if (relu) {
    if (shape == 331) SB_POOL_WG(true, 331); else if (shape == 333) SB_POOL_WG(true, 333); else SB_POOL_WG(true, 551);
  } else {
    if (shape == 331) SB_POOL_WG(false, 331); else if (shape == 333) SB_POOL_WG(false, 333); else SB_POOL_WG(false, 551);
  }
#undef SB_POOL_WG
  SB_LAUNCHED();
  if (defer && defer->add(ws, dw, K * g.C, blocks)) return true;
  partial_reduce(ws, dw, K * g.C, blocks, s);
  return true;
}
bool pool22_map_ok(const PoolGeom& g) {
  return pool_is_22s1(g) && (g.C & 3) == 0 && (long long)g.N * g.H * 8 * (g.C / 4) < (1LL << 31);
}
void pool_max_bwd_v4(const float* x, const float* dy, float* dx, const PoolGeom& g, int relu, float thr, cudaStream_t s,
                     const unsigned char* map) {
  if (pool_is_22s1(g) && (long long)g.N * g.H * 8 * (g.C / 4) < (1LL << 31)) {  // <= 8 segments; 32-bit thread index in the kernel
    int segs, seg_len;
    pool22_segments(g, 5, &segs, &seg_len);
    const long long nt = (long long)g.N * g.H * segs * (g.C / 4);
    PoolWgradFuse none{};
    if (map) {
      launch_k(pool22_bwd_map_kernel<0>, (int)((nt + 127) / 128), 128, 0, s, (const unsigned*)map, dy, dx, g, relu, seg_len, segs, nt, none);
      SB_LAUNCHED();
      return;
    }
    if (relu) launch_k(pool22_bwd_strip_kernel<true, 0>, (int)((nt + 127) / 128), 128, 0, s, x, dy, dx, g, thr, seg_len, segs, nt, none);
    else launch_k(pool22_bwd_strip_kernel<false, 0>, (int)((nt + 127) / 128), 128, 0, s, x, dy, dx, g, 0.f, seg_len, segs, nt, none);
    SB_LAUNCHED();
    return;
  }
  const long long n4 = (long long)g.N * g.H * g.W * (g.C / 4);
  if (relu) launch_k(pool_max_bwd_v4_kernel<true>, grid_for(n4, 256), 256, 0, s, x, dy, dx, g, thr, n4);
  else launch_k(pool_max_bwd_v4_kernel<false>, grid_for(n4, 256), 256, 0, s, x, dy, dx, g, 0.f, n4);
  SB_LAUNCHED();
}

// =====================================================================================================================
// Conv2D with a tiny reduction: K = KH*KW*Cin <= 32, Cout % 4 == 0 (any stride / padding)
// =====================================================================================================================
constexpr int kSmallK = 32;
__global__ void __launch_bounds__(256) conv_smallk_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              float* __restrict__ y, ConvGeom g, int relu, float thr, long long n4) {
  pdl_prologue();
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
// Specialised forward for the filter shapes the first layer of the BASELINE CNNs uses (stride 1): a thread owns 4 output
// channels x PW consecutive output pixels of one row, so the input patch and the filter slab (shared memory) are each read once
// for PW pixels and every loop is fully unrolled.
template <int KH, int KW, int CIN, int PW>
__global__ void __launch_bounds__(256) conv_smallk_fwd_tile_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   float* __restrict__ y, ConvGeom g, int relu, float thr,
                                                                   int segs, long long n_threads) {
  pdl_prologue();
  constexpr int K = KH * KW * CIN;
  const int C4 = g.Cout >> 2;
  extern __shared__ __align__(16) float swt[];  // filter [K][Cout]
  for (int i = threadIdx.x; i < K * g.Cout; i += blockDim.x) swt[i] = w[i];
  __syncthreads();
  // persistent: the filter is staged once per block, the items advance by the grid's thread count (32-bit index split: the launcher
  // keeps n_threads < 2^31).  Packed fp32x2 FMAs (two independent IEEE FMAs, the same bits as the scalar form): channel pairs
  // (0,1) and (2,3), as in conv_smallk_pool_fwd_kernel below.
  const unsigned nthr = (unsigned)n_threads, step = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nthr; i += step) {
    const int c4 = (int)(i % (unsigned)C4);
    unsigned t = i / (unsigned)C4;
    const int seg = (int)(t % (unsigned)segs); t /= (unsigned)segs;
    const int oh = (int)(t % (unsigned)g.OH);
    const int b = (int)(t / (unsigned)g.OH);
    const int ow0 = seg * PW;
    float2 acc[PW][2];
#pragma unroll
    for (int p = 0; p < PW; ++p) acc[p][0] = acc[p][1] = make_float2(0.f, 0.f);
    const float* wq = swt + c4 * 4;
#pragma unroll
    for (int a = 0; a < KH; ++a) {
      const int ih = oh + a - g.PT;
      const bool row_ok = ih >= 0 && ih < g.H;
      const float* xr = x + ((long long)b * g.H + (row_ok ? ih : 0)) * g.W * CIN;
      float xv[(PW + KW - 1) * CIN];
#pragma unroll
      for (int q = 0; q < PW + KW - 1; ++q) {
        const int iw = ow0 + q - g.PL;
        const bool ok = row_ok && iw >= 0 && iw < g.W;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) xv[q * CIN + ci] = ok ? __ldg(xr + iw * CIN + ci) : 0.f;
      }
#pragma unroll
      for (int bb = 0; bb < KW; ++bb)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float4 wv = *reinterpret_cast<const float4*>(wq + ((a * KW + bb) * CIN + ci) * g.Cout);
          const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
#pragma unroll
          for (int p = 0; p < PW; ++p) {
            const float v = xv[(p + bb) * CIN + ci];
            const float2 vv = make_float2(v, v);
            acc[p][0] = __ffma2_rn(vv, w01, acc[p][0]);
            acc[p][1] = __ffma2_rn(vv, w23, acc[p][1]);
          }
        }
    }
    float* yo = y + (((long long)b * g.OH + oh) * g.OW + ow0) * g.Cout + c4 * 4;
#pragma unroll
    for (int p = 0; p < PW; ++p) {
      if (ow0 + p >= g.OW) break;
      float4 o = make_float4(acc[p][0].x, acc[p][0].y, acc[p][1].x, acc[p][1].y);
      if (relu) { o.x = fmaxf(thr, o.x); o.y = fmaxf(thr, o.y); o.z = fmaxf(thr, o.z); o.w = fmaxf(thr, o.w); }
      st4(yo + (long long)p * g.Cout, o);
    }
  }
}
template <int KH, int KW, int CIN>
static bool launch_smallk_tile(const float* x, const float* w, float* y, const ConvGeom& g, int relu, float thr, cudaStream_t s) {
  constexpr int PW = 4;
  const int segs = (g.OW + PW - 1) / PW;
  const long long nt = (long long)g.N * g.OH * segs * (g.Cout / 4);
  if (nt >= (1ll << 31) || (long long)g.H * g.W * CIN >= (1ll << 31)) return false;  // 32-bit item index in the kernel
  const size_t smem = (size_t)KH * KW * CIN * g.Cout * sizeof(float);
  const int blocks = (int)std::min<long long>((nt + 255) / 256, (long long)kNumSMs * 8);
  launch_k(conv_smallk_fwd_tile_kernel<KH, KW, CIN, PW>, blocks, 256, smem, s, x, w, y, g, relu, thr, segs, nt);
  SB_LAUNCHED();
  return true;
}
// Small-K conv (stride 1) >> [ReLU] >> max-pool 2x2 stride 1 VALID in ONE pass: a thread owns 4 channels x PW pool outputs of pool
// row pr, i.e. the conv outputs of rows pr, pr+1 and columns ow0 .. ow0+PW.  It stores its own conv row (the backward needs the
// activations: ReLU mask and pool arg-max) and the pooled row; conv row pr+1 is recomputed by the thread below with the same FMA
// order, so the stored activation and the pooled maximum are the same bits.  The conv output is never re-read.
// MAP: instead of the conv activations the kernel stores one winner byte per pool output (win4) -- everything the backward pass
// needs from them -- so the conv output is never written at all.
template <int KH, int KW, int CIN, int PW, bool MAP>
__global__ void __launch_bounds__(256) conv_smallk_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   float* __restrict__ y, float* __restrict__ pooled, ConvGeom g,
                                                                   int relu, float thr, int segs, long long n_threads,
                                                                   int prewait, unsigned* __restrict__ map) {
  constexpr int K = KH * KW * CIN;
  const int C4 = g.Cout >> 2;
  extern __shared__ __align__(16) float swt[];  // filter [K][Cout]
  // prewait (common.cuh: prewait_ok): the kernel launched just before this one does not write weights, so the filter is staged
  // under that kernel's tail, before the dependency wait; otherwise after it
  if (!prewait) pdl_wait();
  for (int i = threadIdx.x; i < K * g.Cout; i += blockDim.x) swt[i] = w[i];
  if (prewait) pdl_wait();
  pdl_trigger();
  __syncthreads();
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;  // 32-bit index split (the launcher keeps n_threads < 2^31)
  if ((long long)i >= n_threads) return;
  const int OHp = g.OH - 1, OWp = g.OW - 1;
  const int c4 = (int)(i % (unsigned)C4);
  unsigned t = i / (unsigned)C4;
  const int seg = (int)(t % (unsigned)segs); t /= (unsigned)segs;
  const int pr = (int)(t % (unsigned)OHp);
  const int b = (int)(t / (unsigned)OHp);
  const int ow0 = seg * PW;
  // packed fp32x2 FMAs (FFMA2: two independent IEEE FMAs, the same bits as the scalar form): channel pairs (0,1) and (2,3)
  float2 acc[2][PW + 1][2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int p = 0; p <= PW; ++p) acc[r][p][0] = acc[r][p][1] = make_float2(0.f, 0.f);
  // single-channel rows without left padding whose pitch keeps 16-byte alignment: the PW + KW <= 8 inputs of a row are two
  // float4 loads (the second only when it stays inside the row) instead of PW + KW bounds-checked scalar loads
  const bool vec_rows = CIN == 1 && PW + KW <= 8 && g.PL == 0 && (g.W & 3) == 0 && (PW & 3) == 0 && ((uintptr_t)x & 15) == 0;
  // (ow0 and W are multiples of 4: each of the two quads is entirely inside or entirely outside the row)
  const float* wq = swt + c4 * 4;
#pragma unroll
  for (int a = 0; a <= KH; ++a) {  // input row pr + a - PT feeds conv row pr (filter row a) and conv row pr+1 (filter row a-1)
    const int ih = pr + a - g.PT;
    const bool row_ok = ih >= 0 && ih < g.H;
    const float* xr = x + ((long long)b * g.H + (row_ok ? ih : 0)) * g.W * CIN;
    float xv[(PW + KW) * CIN > 8 ? (PW + KW) * CIN : 8];
    if (vec_rows) {
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 lo = (row_ok && ow0 + 4 <= g.W) ? ld4(xr + ow0) : z4;  // outside only with explicit right padding (OW > W)
      const float4 hi = (row_ok && ow0 + 8 <= g.W) ? ld4(xr + ow0 + 4) : z4;
      xv[0] = lo.x; xv[1] = lo.y; xv[2] = lo.z; xv[3] = lo.w; xv[4] = hi.x; xv[5] = hi.y; xv[6] = hi.z; xv[7] = hi.w;
    } else {
#pragma unroll
      for (int q = 0; q < PW + KW; ++q) {
        const int iw = ow0 + q - g.PL;
        const bool ok = row_ok && iw >= 0 && iw < g.W;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) xv[q * CIN + ci] = ok ? __ldg(xr + iw * CIN + ci) : 0.f;
      }
    }
#pragma unroll
    for (int bb = 0; bb < KW; ++bb)
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        if (a < KH) {
          const float4 wv = *reinterpret_cast<const float4*>(wq + ((a * KW + bb) * CIN + ci) * g.Cout);
          const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
#pragma unroll
          for (int p = 0; p <= PW; ++p) {
            const float v = xv[(p + bb) * CIN + ci];
            const float2 vv = make_float2(v, v);
            acc[0][p][0] = __ffma2_rn(vv, w01, acc[0][p][0]);
            acc[0][p][1] = __ffma2_rn(vv, w23, acc[0][p][1]);
          }
        }
        if (a >= 1) {
          const float4 wv = *reinterpret_cast<const float4*>(wq + (((a - 1) * KW + bb) * CIN + ci) * g.Cout);
          const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
#pragma unroll
          for (int p = 0; p <= PW; ++p) {
            const float v = xv[(p + bb) * CIN + ci];
            const float2 vv = make_float2(v, v);
            acc[1][p][0] = __ffma2_rn(vv, w01, acc[1][p][0]);
            acc[1][p][1] = __ffma2_rn(vv, w23, acc[1][p][1]);
          }
        }
      }
  }
  float4 out[2][PW + 1];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int p = 0; p <= PW; ++p) {
      float4 o = make_float4(acc[r][p][0].x, acc[r][p][0].y, acc[r][p][1].x, acc[r][p][1].y);
      if (relu) { o.x = fmaxf(thr, o.x); o.y = fmaxf(thr, o.y); o.z = fmaxf(thr, o.z); o.w = fmaxf(thr, o.w); }
      out[r][p] = o;
    }
  float* po = pooled + (((long long)b * OHp + pr) * OWp + ow0) * g.Cout + c4 * 4;
  if (MAP) {
    unsigned* mo = map + (((long long)b * OHp + pr) * OWp + ow0) * C4 + c4;  // one byte per channel: a 32-bit word per quad
#pragma unroll
    for (int p = 0; p < PW; ++p) {
      if (ow0 + p >= OWp) break;
      float4 m;
      const unsigned b0 = win4(out[0][p].x, out[0][p + 1].x, out[1][p].x, out[1][p + 1].x, relu, thr, &m.x);
      const unsigned b1 = win4(out[0][p].y, out[0][p + 1].y, out[1][p].y, out[1][p + 1].y, relu, thr, &m.y);
      const unsigned b2 = win4(out[0][p].z, out[0][p + 1].z, out[1][p].z, out[1][p + 1].z, relu, thr, &m.z);
      const unsigned b3 = win4(out[0][p].w, out[0][p + 1].w, out[1][p].w, out[1][p + 1].w, relu, thr, &m.w);
      st4(po + p * g.Cout, m);
      mo[p * C4] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
    return;
  }
  // the conv activations: own row, and the last conv row from the last pool row
  const int rows_out = pr == OHp - 1 ? 2 : 1;
  for (int r = 0; r < rows_out; ++r) {
    float* yo = y + (((long long)b * g.OH + pr + r) * g.OW + ow0) * g.Cout + c4 * 4;
#pragma unroll
    for (int p = 0; p < PW; ++p)
      if (ow0 + p < g.OW) st4(yo + p * g.Cout, out[r][p]);
  }
#pragma unroll
  for (int p = 0; p < PW; ++p) {
    if (ow0 + p >= OWp) break;
    float4 m;
    m.x = fmaxf(fmaxf(out[0][p].x, out[0][p + 1].x), fmaxf(out[1][p].x, out[1][p + 1].x));
    m.y = fmaxf(fmaxf(out[0][p].y, out[0][p + 1].y), fmaxf(out[1][p].y, out[1][p + 1].y));
    m.z = fmaxf(fmaxf(out[0][p].z, out[0][p + 1].z), fmaxf(out[1][p].z, out[1][p + 1].z));
    m.w = fmaxf(fmaxf(out[0][p].w, out[0][p + 1].w), fmaxf(out[1][p].w, out[1][p + 1].w));
    st4(po + p * g.Cout, m);
  }
}
template <int KH, int KW, int CIN>
static void launch_smallk_pool(const float* x, const float* w, float* y, float* pooled, const ConvGeom& g, int relu, float thr,
                               unsigned char* map, cudaStream_t s) {
  constexpr int PW = 4;
  const int segs = (g.OW + PW - 1) / PW;  // covers the conv columns; the pool columns are one fewer
  const long long nt = (long long)g.N * (g.OH - 1) * segs * (g.Cout / 4);
  const size_t smem = (size_t)KH * KW * CIN * g.Cout * sizeof(float);
  if (map)
    launch_k(conv_smallk_pool_fwd_kernel<KH, KW, CIN, PW, true>, (int)((nt + 255) / 256), 256, smem, s, x, w, y, pooled, g, relu, thr,
             segs, nt, prewait_ok(), (unsigned*)map);
  else
    launch_k(conv_smallk_pool_fwd_kernel<KH, KW, CIN, PW, false>, (int)((nt + 255) / 256), 256, smem, s, x, w, y, pooled, g, relu, thr,
             segs, nt, prewait_ok(), (unsigned*)nullptr);
  SB_LAUNCHED();
}
bool conv_smallk_pool_fwd(const float* x, const float* w, float* y, float* pooled, const ConvGeom& g, int relu, float thr,
                          unsigned char* map, cudaStream_t s) {
  if (g.SH != 1 || g.SW != 1 || g.OH < 2 || g.OW < 2 || g.Cout % 4) return false;
  if ((long long)g.N * g.OH * g.OW * g.Cout >= (1LL << 31)) return false;  // 32-bit thread index and in-row offsets in the kernel
  const int shape = g.KH * 100 + g.KW * 10 + g.Cin;
  if (shape == 331) { launch_smallk_pool<3, 3, 1>(x, w, y, pooled, g, relu, thr, map, s); return true; }
  if (shape == 333) { launch_smallk_pool<3, 3, 3>(x, w, y, pooled, g, relu, thr, map, s); return true; }
  return false;
}

void conv_smallk_fwd(const float* x, const float* w, float* y, const ConvGeom& g, int relu, float thr, cudaStream_t s) {
  if (g.SH == 1 && g.SW == 1) {
    const int shape = g.KH * 100 + g.KW * 10 + g.Cin;
    if (shape == 331 && launch_smallk_tile<3, 3, 1>(x, w, y, g, relu, thr, s)) return;
    if (shape == 333 && launch_smallk_tile<3, 3, 3>(x, w, y, g, relu, thr, s)) return;
    if (shape == 551 && launch_smallk_tile<5, 5, 1>(x, w, y, g, relu, thr, s)) return;
  }
  const long long n4 = (long long)g.N * g.OH * g.OW * (g.Cout / 4);
  launch_k(conv_smallk_fwd_kernel, grid_for(n4, 256), 256, 0, s, x, w, y, g, relu, thr, n4);
  SB_LAUNCHED();
}

// wgrad: dW[k][co] = sum_pixels X(pixel, k) * dY[pixel][co].  Block = (Cout/4) channel groups x PL pixel lanes; every thread
// keeps K float4 accumulators; lanes are tree-free summed in lane order through shared memory; one partial per block.
template <int KH, int KW, int CIN>
__global__ void __launch_bounds__(256) conv_smallk_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                float* __restrict__ partial, ConvGeom g, long long pixels,
                                                                long long pix_per_block) {
  pdl_prologue();
  constexpr int K = KH * KW * CIN;
  const int C4 = g.Cout >> 2;
  const int lanes = blockDim.x / C4;
  const int c4 = threadIdx.x % C4, lane = threadIdx.x / C4;
  float4 acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  // every pixel lane walks a CONTIGUOUS run of output pixels: (b, oh, ow) advance incrementally (no divisions in the loop),
  // neighbouring pixels share input taps in L1, and the next pixel's dY is fetched while this one is accumulated
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  const long long p1 = p0 + pix_per_block < pixels ? p0 + pix_per_block : pixels;
  const long long run = (p1 - p0 + lanes - 1) / lanes;
  long long p = p0 + lane * run;
  const long long pe = p + run < p1 ? p + run : p1;
  int ow = (int)(p % g.OW);
  int oh = (int)((p / g.OW) % g.OH);
  int b = (int)(p / ((long long)g.OW * g.OH));
  float4 gy = p < pe ? ld4(dy + p * g.Cout + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (; p < pe; ++p) {
    const float4 gcur = gy;
    if (p + 1 < pe) gy = ld4(dy + (p + 1) * g.Cout + c4 * 4);
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
          r.x = fmaf(xv, gcur.x, r.x); r.y = fmaf(xv, gcur.y, r.y); r.z = fmaf(xv, gcur.z, r.z); r.w = fmaf(xv, gcur.w, r.w);
        }
      }
    }
    if (++ow == g.OW) {
      ow = 0;
      if (++oh == g.OH) { oh = 0; ++b; }
    }
  }
  // cross-lane sum, fixed order: (1) the blockDim/32 ... pixel lanes that share a warp by shuffles (xor C4, 2*C4, ..),
  // (2) the warps in warp order through shared memory.  Requires C4 to be a power of two <= 32 (checked by the launcher).
  const int wl = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    for (int o = C4; o < 32; o <<= 1) {
      acc[k].x = __fadd_rn(acc[k].x, __shfl_xor_sync(0xffffffffu, acc[k].x, o));
      acc[k].y = __fadd_rn(acc[k].y, __shfl_xor_sync(0xffffffffu, acc[k].y, o));
      acc[k].z = __fadd_rn(acc[k].z, __shfl_xor_sync(0xffffffffu, acc[k].z, o));
      acc[k].w = __fadd_rn(acc[k].w, __shfl_xor_sync(0xffffffffu, acc[k].w, o));
    }
  }
  extern __shared__ __align__(16) float4 red[];  // [nwarps][K][C4]
  if (wl < C4) {
#pragma unroll
    for (int k = 0; k < K; ++k) red[(wid * K + k) * C4 + wl] = acc[k];
  }
  __syncthreads();
  float* part = partial + (long long)blockIdx.x * K * g.Cout;
  for (int i = threadIdx.x; i < K * C4; i += blockDim.x) {
    float4 tot = red[i];
    for (int wq = 1; wq < nwarps; ++wq) {
      const float4 v = red[wq * K * C4 + i];
      tot.x = __fadd_rn(tot.x, v.x); tot.y = __fadd_rn(tot.y, v.y); tot.z = __fadd_rn(tot.z, v.z); tot.w = __fadd_rn(tot.w, v.w);
    }
    st4(part + (i / C4) * g.Cout + (i % C4) * 4, tot);
  }
}
void partial_reduce(const float* partial, float* out, int n, int parts, cudaStream_t s);
// wgrad through shared memory (Cout >= 32): a block walks tiles of 64 consecutive output pixels.  Per tile the dY rows [64][Cout] and
// the tile's im2col matrix [64][K padded to 32] are staged once; thread (phase, k group of 8, channel group of 4) then accumulates
// 8 x 4 products per pixel from three 128-bit shared loads (the two X loads are broadcasts), i.e. 32 FMAs per 3 loads instead of
// 108 FMAs per 28 global loads above -- and 32 accumulators instead of 108, so three blocks fit an SM.  The phases (pixels
// p % phases) are summed in phase order, one partial per block, blocks reduced in block order (deterministic).
constexpr int kWgP = 64;        // pixels per tile
constexpr int kWgKP = 32;       // K padded
constexpr int kWgColStride = 36;  // floats per im2col row: 16-byte aligned rows, 4-way (not 32-way) conflicts on the scalar fill
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool live) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = live ? 16 : 0;  // 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int KH, int KW, int CIN, int Q>
__device__ __forceinline__ void wg_load_cols(float (&v)[8], const float* __restrict__ x, const ConvGeom& g, int b, int oh, int ow, bool live) {
  constexpr int K = KH * KW * CIN;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = Q * 8 + j;
    v[j] = 0.f;
    if (k < K) {
      const int tap = k / CIN, ci = k % CIN, a = tap / KW, bb = tap % KW;
      const int ih = oh * g.SH + a - g.PT, iw = ow * g.SW + bb - g.PL;
      if (live && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) v[j] = __ldg(x + (((long long)b * g.H + ih) * g.W + iw) * CIN + ci);
    }
  }
}
// Two tile buffers: while the block accumulates tile t, the dY rows of tile t+1 arrive through cp.async and its im2col values sit in
// registers (loaded before the FMAs, stored after them) -- one barrier per tile, no exposed global latency.
template <int KH, int KW, int CIN>
__global__ void __launch_bounds__(256, 2) conv_smallk_wgrad_tile_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                        float* __restrict__ partial, ConvGeom g, long long pixels, int tiles,
                                                                        int buf_floats) {
  pdl_prologue();
  constexpr int K = KH * KW * CIN;
  static_assert(K <= kWgKP, "K must fit the padded im2col row");
  extern __shared__ __align__(16) float wg_sm[];
  const int Cout = g.Cout, CO4 = Cout >> 2;
  const int tid = threadIdx.x;
  const int tpp = CO4 * (kWgKP / 8);         // threads per phase
  const int phases = 256 / tpp;
  const int ph = tid / tpp, r = tid % tpp, kg = r / CO4, c4 = r % CO4;
  float2 acc[8][2];  // packed fp32x2 FMAs (two independent IEEE FMAs): channel pairs (0,1) and (2,3)
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
  const int fp = tid % kWgP, fq = tid / kWgP;  // im2col fill: pixel of the tile, quarter of the padded K (warp-uniform)
  // buffer b: dY rows [kWgP][Cout] then im2col rows [kWgP][kWgColStride]
  auto issue_dy = [&](int tile, float* dys) {
    const long long p0 = (long long)tile * kWgP;
    for (int i = tid; i < kWgP * CO4; i += 256) {
      const bool live = p0 + i / CO4 < pixels;
      cp_async16(dys + i * 4, dy + (live ? p0 * Cout + (long long)i * 4 : 0), live);
    }
  };
  auto load_cols = [&](int tile, float (&v)[8]) {
    const unsigned p = (unsigned)tile * kWgP + fp;  // the launcher keeps pixels < 2^31
    const bool live = p < (unsigned)pixels;
    const unsigned pr = p / (unsigned)g.OW;
    const int ow = (int)(p - pr * (unsigned)g.OW), b = (int)(pr / (unsigned)g.OH), oh = (int)(pr - (unsigned)b * (unsigned)g.OH);
    if (fq == 0) wg_load_cols<KH, KW, CIN, 0>(v, x, g, b, oh, ow, live);
    else if (fq == 1) wg_load_cols<KH, KW, CIN, 1>(v, x, g, b, oh, ow, live);
    else if (fq == 2) wg_load_cols<KH, KW, CIN, 2>(v, x, g, b, oh, ow, live);
    else wg_load_cols<KH, KW, CIN, 3>(v, x, g, b, oh, ow, live);
  };
  auto store_cols = [&](float* cols, const float (&v)[8]) {
    float* col = cols + fp * kWgColStride + fq * 8;
    *reinterpret_cast<float4*>(col) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(col + 4) = make_float4(v[4], v[5], v[6], v[7]);
  };
  int tile = blockIdx.x, cur = 0;
  if (tile < tiles) {
    float v[8];
    issue_dy(tile, wg_sm);
    load_cols(tile, v);
    store_cols(wg_sm + kWgP * Cout, v);
  }
  for (; tile < tiles; tile += gridDim.x, cur ^= 1) {
    float* dys = wg_sm + (size_t)cur * buf_floats;
    float* cols = dys + kWgP * Cout;
    float* ndys = wg_sm + (size_t)(cur ^ 1) * buf_floats;
    cp_async_wait_all();
    __syncthreads();  // tile `cur` is complete; everyone has left the other buffer (read during the previous iteration)
    const int next = tile + gridDim.x;
    float v[8];
    if (next < tiles) {
      issue_dy(next, ndys);
      load_cols(next, v);
    }
    if (ph < phases) {
#pragma unroll 4
      for (int p = ph; p < kWgP; p += phases) {
        const float4 xa = *reinterpret_cast<const float4*>(cols + p * kWgColStride + kg * 8);
        const float4 xb = *reinterpret_cast<const float4*>(cols + p * kWgColStride + kg * 8 + 4);
        const float4 gv = *reinterpret_cast<const float4*>(dys + p * Cout + c4 * 4);
        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const float2 g01 = make_float2(gv.x, gv.y), g23 = make_float2(gv.z, gv.w);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 vv = make_float2(xv[i], xv[i]);
          acc[i][0] = __ffma2_rn(vv, g01, acc[i][0]);
          acc[i][1] = __ffma2_rn(vv, g23, acc[i][1]);
        }
      }
    }
    if (next < tiles) store_cols(ndys + kWgP * Cout, v);
  }
  cp_async_wait_all();
  __syncthreads();
  // the phases, in phase order: red[phase][k][Cout] reuses the tile buffers
  float* red = wg_sm;
  if (ph < phases) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(red + ((size_t)ph * kWgKP + kg * 8 + i) * Cout + c4 * 4) = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
  }
  __syncthreads();
  float* part = partial + (long long)blockIdx.x * K * Cout;
  for (int i = tid; i < K * CO4; i += 256) {
    const int k = i / CO4, cc = i % CO4;
    float4 tot = *reinterpret_cast<const float4*>(red + (size_t)k * Cout + cc * 4);
    for (int q = 1; q < phases; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(red + ((size_t)q * kWgKP + k) * Cout + cc * 4);
      tot.x = __fadd_rn(tot.x, v.x); tot.y = __fadd_rn(tot.y, v.y); tot.z = __fadd_rn(tot.z, v.z); tot.w = __fadd_rn(tot.w, v.w);
    }
    st4(part + (size_t)k * Cout + cc * 4, tot);
  }
}
template <int KH, int KW, int CIN>
static bool launch_smallk_wgrad_tile(const float* x, const float* dy, float* dw, const ConvGeom& g, float* ws, size_t ws_floats, cudaStream_t s) {
  const int K = KH * KW * CIN, CO4 = g.Cout / 4;
  const int tpp = CO4 * (kWgKP / 8);
  if (CO4 < 8 || (CO4 & (CO4 - 1)) || tpp > 256) return false;  // Cout = 32, 64, 128, 256
  const int phases = 256 / tpp;
  const size_t tile_floats = (size_t)kWgP * g.Cout + (size_t)kWgP * kWgColStride;
  const size_t red_floats = (size_t)phases * kWgKP * g.Cout;
  const size_t smem = std::max(2 * tile_floats, red_floats) * sizeof(float);
  if (smem > 200 * 1024) return false;
  const long long pixels = (long long)g.N * g.OH * g.OW;
  if (pixels + kWgP >= (1ll << 31)) return false;  // 32-bit pixel index in the kernel
  const int tiles = (int)((pixels + kWgP - 1) / kWgP);
  int blocks = std::min(tiles, 2 * kNumSMs);
  if ((size_t)blocks * K * g.Cout > ws_floats) blocks = (int)(ws_floats / ((size_t)K * g.Cout));
  if (blocks < 1) return false;
  static bool attr = false;
  if (!attr) {
    SB_CUDA(cudaFuncSetAttribute(conv_smallk_wgrad_tile_kernel<KH, KW, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  launch_k(conv_smallk_wgrad_tile_kernel<KH, KW, CIN>, blocks, 256, smem, s, x, dy, ws, g, pixels, tiles, (int)tile_floats);
  SB_LAUNCHED();
  partial_reduce(ws, dw, K * g.Cout, blocks, s);
  return true;
}

// filter shapes 3x3x1, 3x3x3, 5x5x1 (fully unrolled accumulators); ws must hold blocks*K*Cout floats
bool conv_smallk_wgrad(const float* x, const float* dy, float* dw, const ConvGeom& g, float* ws, size_t ws_floats, cudaStream_t s) {
  const int K = g.KH * g.KW * g.Cin;
  const int C4 = g.Cout / 4;
  const int shape = g.KH * 100 + g.KW * 10 + g.Cin;
  if ((shape != 331 && shape != 333 && shape != 551) || C4 > 32 || (C4 & (C4 - 1))) return false;
  {
    const char* old = getenv("SB_SMALLK_WGRAD_REG");  // =1: the register-accumulator kernel for every shape (A/B)
    if (!(old && old[0] == '1') && g.Cout >= 32) {
      if (shape == 331 && launch_smallk_wgrad_tile<3, 3, 1>(x, dy, dw, g, ws, ws_floats, s)) return true;
      if (shape == 333 && launch_smallk_wgrad_tile<3, 3, 3>(x, dy, dw, g, ws, ws_floats, s)) return true;
      if (shape == 551 && launch_smallk_wgrad_tile<5, 5, 1>(x, dy, dw, g, ws, ws_floats, s)) return true;
    }
  }
  const long long pixels = (long long)g.N * g.OH * g.OW;
  const size_t smem = (size_t)8 * K * C4 * sizeof(float4);
  if (smem > 160 * 1024) return false;
  static bool attr = false;
  if (!attr) {
    SB_CUDA(cudaFuncSetAttribute(conv_smallk_wgrad_kernel<3, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    SB_CUDA(cudaFuncSetAttribute(conv_smallk_wgrad_kernel<3, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    SB_CUDA(cudaFuncSetAttribute(conv_smallk_wgrad_kernel<5, 5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  int blocks = 2 * kNumSMs;
  if ((size_t)blocks * K * g.Cout > ws_floats) blocks = (int)(ws_floats / ((size_t)K * g.Cout));
  if (blocks < 1) return false;
  if (blocks > pixels) blocks = (int)pixels;
  const long long ppb = (pixels + blocks - 1) / blocks;
  blocks = (int)((pixels + ppb - 1) / ppb);
  if (shape == 331) launch_k(conv_smallk_wgrad_kernel<3, 3, 1>, blocks, 256, smem, s, x, dy, ws, g, pixels, ppb);
  else if (shape == 333) launch_k(conv_smallk_wgrad_kernel<3, 3, 3>, blocks, 256, smem, s, x, dy, ws, g, pixels, ppb);
  else launch_k(conv_smallk_wgrad_kernel<5, 5, 1>, blocks, 256, smem, s, x, dy, ws, g, pixels, ppb);
  SB_LAUNCHED();
  const int n = K * g.Cout;
  partial_reduce(ws, dw, n, blocks, s);
  return true;
}

// =====================================================================================================================
// skinny dense head: Z[M,N] = X[M,K] . W[K,N] with N <= 16 outputs, M <= 128 rows and a long reduction K
// (layer/Dense.scala:57-61 feeding Bias + Softmax + CategoricalCrossentropy).  Everything here is bound by reading X
// (forward, wgrad) or writing dX (dgrad) once.
// =====================================================================================================================
constexpr int kSkinnyN = 16;
constexpr int kHeadThreads = 128;

// generic deterministic column reduction  out[i] = sum_c partial[c][i]  (c ascending inside a lane group, then a fixed tree)
template <int YG>
__global__ void __launch_bounds__(32 * YG) partial_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int n,
                                                                int parts, int early) {
  pdl_prologue_trigger_if(early);
  __shared__ float red[YG][33];
  const int i = blockIdx.x * 32 + threadIdx.x, y = threadIdx.y;
  float acc = 0.f;
  if (i < n) {
#pragma unroll 4
    for (int c = y; c < parts; c += YG) acc = __fadd_rn(acc, __ldg(partial + (size_t)c * n + i));
  }
  red[y][threadIdx.x] = acc;
  __syncthreads();
  if (y == 0 && i < n) {
    float t = red[0][threadIdx.x];
#pragma unroll
    for (int j = 1; j < YG; ++j) t = __fadd_rn(t, red[j][threadIdx.x]);
    out[i] = t;
  }
}
// the same sums (same order per element, so the same bits) four elements per thread: the big filter gradients (n >= 16 K floats) are
// bound by reading the partials, and 128-bit loads of 512 contiguous bytes per warp do that at several times the scalar kernel's rate
template <int YG>
__global__ void __launch_bounds__(32 * YG) partial_reduce4_kernel(const float* __restrict__ partial, float* __restrict__ out, int n4,
                                                                 int parts, int early) {
  pdl_prologue_trigger_if(early);
  __shared__ float4 red[YG][33];
  const int i = blockIdx.x * 32 + threadIdx.x, y = threadIdx.y;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
    const float4* p4 = reinterpret_cast<const float4*>(partial) + i;
#pragma unroll 4
    for (int c = y; c < parts; c += YG) {
      const float4 v = __ldg(p4 + (size_t)c * n4);
      acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
    }
  }
  red[y][threadIdx.x] = acc;
  __syncthreads();
  if (y == 0 && i < n4) {
    float4 t = red[0][threadIdx.x];
#pragma unroll
    for (int j = 1; j < YG; ++j) {
      const float4 v = red[j][threadIdx.x];
      t.x = __fadd_rn(t.x, v.x); t.y = __fadd_rn(t.y, v.y); t.z = __fadd_rn(t.z, v.z); t.w = __fadd_rn(t.w, v.w);
    }
    reinterpret_cast<float4*>(out)[i] = t;
  }
}
// several reductions in ONE launch (the filter-gradient partials of different layers, all due just before the optimizer): block b
// belongs to the segment whose block range contains it; each segment is summed exactly as partial_reduce_kernel<32> would
__global__ void __launch_bounds__(32 * 32) partial_reduce_multi_kernel(ReduceBatch rb, int early) {
  pdl_prologue_trigger_if(early);
  __shared__ float red[32][33];
  int seg = 0, b = blockIdx.x;
  while (seg + 1 < rb.count && b >= rb.blocks[seg]) { b -= rb.blocks[seg]; ++seg; }
  const float* __restrict__ partial = rb.partial[seg];
  float* __restrict__ out = rb.out[seg];
  const int n = rb.n[seg], parts = rb.parts[seg];
  const int i = b * 32 + threadIdx.x, y = threadIdx.y;
  float acc = 0.f;
  if (i < n)
    for (int c = y; c < parts; c += 32) acc = __fadd_rn(acc, __ldg(partial + (size_t)c * n + i));
  red[y][threadIdx.x] = acc;
  __syncthreads();
  if (y == 0 && i < n) {
    float t = red[0][threadIdx.x];
#pragma unroll
    for (int j = 1; j < 32; ++j) t = __fadd_rn(t, red[j][threadIdx.x]);
    out[i] = t;
  }
}
void partial_reduce_multi(ReduceBatch& rb, cudaStream_t s) {
  if (rb.count == 0) return;
  int total = 0;
  for (int i = 0; i < rb.count; ++i) {
    rb.blocks[i] = (rb.n[i] + 31) / 32;
    total += rb.blocks[i];
  }
  launch_k(partial_reduce_multi_kernel, total, dim3(32, 32), 0, s, rb, pdl_early_hint());
  SB_LAUNCHED();
  rb.count = 0;
}
// part c goes to lane group c % YG (summed in ascending c inside the group), then the groups in order: a fixed order for a
// given `parts`.  Many partials -> 32 groups so that no thread walks more than ~parts/32 dependent loads.
void partial_reduce(const float* partial, float* out, int n, int parts, cudaStream_t s) {
  if (n >= 16384 && n % 4 == 0 && (((uintptr_t)partial | (uintptr_t)out) & 15) == 0) {
    const int n4 = n / 4;
    if (parts > 48) launch_k(partial_reduce4_kernel<32>, (n4 + 31) / 32, dim3(32, 32), 0, s, partial, out, n4, parts, pdl_early_hint());
    else launch_k(partial_reduce4_kernel<8>, (n4 + 31) / 32, dim3(32, 8), 0, s, partial, out, n4, parts, pdl_early_hint());
    SB_LAUNCHED();
    return;
  }
  if (parts > 48) launch_k(partial_reduce_kernel<32>, (n + 31) / 32, dim3(32, 32), 0, s, partial, out, n, parts, pdl_early_hint());
  else launch_k(partial_reduce_kernel<8>, (n + 31) / 32, dim3(32, 8), 0, s, partial, out, n, parts, pdl_early_hint());
  SB_LAUNCHED();
}

// Forward split-K: CTA c owns the K range [c*KC, c*KC+KC).  The X tile [M][KC] arrives through one bulk async copy per row
// (cp.async.bulk -> mbarrier, all rows in flight at once); the W slab [KC][NT] is staged by the threads meanwhile.  Thread
// (m, kq) then owns output row m over a quarter of the chunk (no cross-lane reduction): per 4 k one conflict-free LDS.128 of X
// (row stride KC+4 floats, (KC+4)/4 odd) and NT/4 broadcast LDS.128 of W per k; the four quarters are summed in order through
// shared memory.  partial[c][m][n].
constexpr int kHeadKQ = 4;
template <int NT>
__global__ void __launch_bounds__(kHeadThreads* kHeadKQ) dense_head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                               float* __restrict__ partial, int Mtot, int K, int N, int KC) {
  pdl_prologue();
  extern __shared__ __align__(128) float sm[];
  const int xs_stride = KC + 4;
  const int Mcap = min(kHeadThreads, Mtot);
  float* xs = sm;                                        // [Mcap][KC+4]
  float* ws = xs + (size_t)Mcap * xs_stride;             // [KC][NT]
  float* red = ws + (size_t)KC * NT;                     // [kHeadKQ][Mcap][NT+1]
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * KC;
  const int kc = min(KC, K - k0);
  // rows [mbase, mbase + M) of the Mtot-row problem (blockIdx.y = row chunk of <= 128 rows)
  const int mbase = blockIdx.y * kHeadThreads;
  const int M = min(kHeadThreads, Mtot - mbase);
  x += (size_t)mbase * K;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) tc::mbar_arrive_expect_tx(&bar, (uint32_t)(M * kc * 4));
  if (tid < M) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     tc::smem_u32(xs + (size_t)tid * xs_stride)),
                 "l"(x + (size_t)tid * K + k0), "r"((uint32_t)(kc * 4)), "r"(tc::smem_u32(&bar))
                 : "memory");
  }
  for (int i = tid; i < KC * NT; i += blockDim.x) {
    const int k = i / NT, n = i - k * NT;
    ws[i] = (k < kc && n < N) ? __ldg(w + (size_t)(k0 + k) * N + n) : 0.f;
  }
  if (kc < KC) {  // tail chunk: the columns the copy does not write must not hold NaN bit patterns (they meet zero weights)
    const int pad = KC - kc;
    for (int i = tid; i < M * pad; i += blockDim.x) xs[(size_t)(i / pad) * xs_stride + kc + i % pad] = 0.f;
  }
  __syncthreads();
  tc::mbar_wait(&bar, 0);
  const int m = tid & (kHeadThreads - 1), kq = tid / kHeadThreads;
  const int KQ = KC / kHeadKQ;
  if (m < M) {
    // packed fp32x2 FMAs (FFMA2): two output columns per instruction
    float2 acc2[NT / 2];
#pragma unroll
    for (int n = 0; n < NT / 2; ++n) acc2[n] = make_float2(0.f, 0.f);
    const float* xr = xs + (size_t)m * xs_stride + kq * KQ;
    const float* wr = ws + (size_t)kq * KQ * NT;
    for (int k = 0; k < KQ; k += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(xr + k);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xx = make_float2(xa[j], xa[j]);
#pragma unroll
        for (int n4 = 0; n4 < NT; n4 += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(wr + (k + j) * NT + n4);
          acc2[n4 / 2] = __ffma2_rn(xx, make_float2(wv.x, wv.y), acc2[n4 / 2]);
          acc2[n4 / 2 + 1] = __ffma2_rn(xx, make_float2(wv.z, wv.w), acc2[n4 / 2 + 1]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NT / 2; ++n) {
      red[((size_t)kq * M + m) * (NT + 1) + 2 * n] = acc2[n].x;
      red[((size_t)kq * M + m) * (NT + 1) + 2 * n + 1] = acc2[n].y;
    }
  }
  __syncthreads();
  for (int i = tid; i < M * N; i += blockDim.x) {
    const int mm = i / N, n = i - mm * N;
    float t = red[(size_t)mm * (NT + 1) + n];
#pragma unroll
    for (int q = 1; q < kHeadKQ; ++q) t = __fadd_rn(t, red[((size_t)q * M + mm) * (NT + 1) + n]);
    partial[((size_t)blockIdx.x * Mtot + mbase) * N + i] = t;
  }
}
static int head_smem_bytes(int M, int KC, int NT) { return (M * (KC + 4) + KC * NT + kHeadKQ * M * (NT + 1)) * 4; }
// writes partial[chunks][M][N] into `partial`; *chunks_out = number of partials.  K % 4 == 0 (16-byte row alignment).
bool dense_head_fwd(const float* x, const float* w, float* partial, int* chunks_out, int M, int N, int K, size_t ws_floats,
                    cudaStream_t s) {
  if (N > kSkinnyN || M < 1 || (K & 3) || (((uintptr_t)x) & 15)) return false;
  const int NT = N <= 12 ? 12 : 16;
  const int m_chunks = (M + kHeadThreads - 1) / kHeadThreads, Mcap = std::min(M, kHeadThreads);
  // CTAs = K chunks x row chunks: one wave when a single row chunk (few partials for the head kernel to sum), else two
  const int want_k = std::max(1, (m_chunks == 1 ? kNumSMs : 2 * kNumSMs) / m_chunks);
  int KC = ((K + want_k - 1) / want_k + 15) & ~15;  // multiple of 16 (=> (KC+4)/4 is odd: conflict-free LDS.128)
  if (KC < 64) KC = 64;
  while (head_smem_bytes(Mcap, KC, NT) > 200 * 1024 && KC > 64) KC = ((KC / 2) + 15) & ~15;
  if (head_smem_bytes(Mcap, KC, NT) > 200 * 1024) return false;
  const int chunks = (K + KC - 1) / KC;
  if ((size_t)chunks * M * N > ws_floats) return false;
  const int smem = head_smem_bytes(Mcap, KC, NT);
  static bool attr = false;
  if (!attr) {
    SB_CUDA(cudaFuncSetAttribute(dense_head_fwd_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SB_CUDA(cudaFuncSetAttribute(dense_head_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const dim3 grid(chunks, m_chunks);
  if (NT == 12) launch_k(dense_head_fwd_kernel<12>, grid, kHeadThreads * kHeadKQ, smem, s, x, w, partial, M, K, N, KC);
  else launch_k(dense_head_fwd_kernel<16>, grid, kHeadThreads * kHeadKQ, smem, s, x, w, partial, M, K, N, KC);
  SB_LAUNCHED();
  *chunks_out = chunks;
  return true;
}
bool dense_skinny_fwd(const float* x, const float* w, float* z, int M, int N, int K, float* ws, size_t ws_floats, cudaStream_t s) {
  int chunks = 0;
  if (!dense_head_fwd(x, w, ws, &chunks, M, N, K, ws_floats, s)) return false;
  partial_reduce(ws, z, M * N, chunks, s);
  return true;
}

// Backward, wgrad and dgrad in ONE pass over X / dX:  dW[k][n] = sum_m X[m][k] G[m][n] ;  dX[m][k] = sum_n G[m][n] W[k][n].
// Thread = one k (coalesced along K) x one of 4 row groups; G and the W slab of the CTA are staged in shared memory; the row
// groups' dW partials are summed in group order through shared memory.
constexpr int kBwdK = 32, kBwdGroups = 4;  // 128-thread CTAs: ~1000 CTAs on cfg2 fill one wave at 7 CTAs/SM
constexpr int kBwdRows = 128;  // rows of G staged per pass
// KPT = k columns per thread: 2 (float2 loads of X / stores of dX, K even) shares every LDS of G between two columns, which
// halves the instructions of the row loop (29 -> 15.5 per row and column)
template <int NT, bool DGRAD, int KPT, int GR = kBwdGroups, int UNR = 8>
__global__ void __launch_bounds__(kBwdK* GR) dense_head_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gmat,
                                                                            const float* __restrict__ w, float* __restrict__ dw,
                                                                            float* __restrict__ dx, int Mtot, int K, int N,
                                                                            int rows_per_split, int early, int prewait) {
  constexpr int KB = kBwdK * KPT;  // k columns of a CTA
  extern __shared__ __align__(128) float sm[];
  // blockIdx.y = row split: rows [mlo, mhi); its dW goes to slice blockIdx.y of `dw` (summed in split order afterwards)
  const int mlo = blockIdx.y * rows_per_split, M = min(Mtot, mlo + rows_per_split);
  dw += (size_t)blockIdx.y * K * N;
  float* sg = sm;                              // [kBwdRows][NT]
  float* sw = sg + (size_t)kBwdRows * NT;      // [KB][NT+1]
  float* red = sw + KB * (NT + 1);             // [GR][KB][NT+1]
  const int tid = threadIdx.x, kk = tid % kBwdK, rg = tid / kBwdK;
  const int k0 = blockIdx.x * KB, k = k0 + kk * KPT;
  const int kn = min(KB, K - k0) * N;  // the CTA's slab of W / dW is contiguous in memory
  const unsigned div_n = (65536u + N - 1) / N;  // i / N for i < 2^16 / N * ... (kn <= 64 * 16): (i * div_n) >> 16
  // prewait (common.cuh: prewait_ok): the kernel before this one (normally the loss kernel) does not write weights, so the W
  // slab is staged before the dependency wait; otherwise after it
  if (!prewait) pdl_wait();
  if (DGRAD) {
    for (int i = tid; i < kn; i += blockDim.x) {
      const int r = (int)(((unsigned)i * div_n) >> 16);
      sw[r * (NT + 1) + (i - r * N)] = __ldg(w + (size_t)k0 * N + i);
    }
  }
  if (prewait) pdl_wait();
  if (early) pdl_trigger();
  __syncthreads();
  float2 acc2[KPT][NT / 2], wr2[KPT][NT / 2];  // packed fp32x2 FMAs (FFMA2)
#pragma unroll
  for (int j = 0; j < KPT; ++j)
#pragma unroll
    for (int n = 0; n < NT / 2; ++n) {
      acc2[j][n] = make_float2(0.f, 0.f);
      const float* swr = sw + (kk * KPT + j) * (NT + 1);
      wr2[j][n] = make_float2((DGRAD && 2 * n < N) ? swr[2 * n] : 0.f, (DGRAD && 2 * n + 1 < N) ? swr[2 * n + 1] : 0.f);
    }
  for (int mb = mlo; mb < M; mb += kBwdRows) {  // row chunks in order: the dW sum runs over rows in ascending chunks
    const int mc = min(kBwdRows, M - mb);
    __syncthreads();
    for (int i = tid; i < mc * NT; i += blockDim.x) {
      const int m = i / NT, n = i - m * NT;
      sg[i] = n < N ? __ldg(gmat + (size_t)(mb + m) * N + n) : 0.f;
    }
    __syncthreads();
    const int rows_per = (mc + GR - 1) / GR;
    const int m0 = rg * rows_per, m1 = min(mc, m0 + rows_per);
    if (k < K) {  // K % KPT == 0 (launcher): the thread's KPT columns are all inside
      const float* xp = x + (size_t)(mb + m0) * K + k;
      float* dxp = dx + (size_t)(mb + m0) * K + k;
#pragma unroll UNR
      for (int m = m0; m < m1; ++m, xp += K, dxp += K) {
        float xv[KPT], dv[KPT];
        if (KPT == 2) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(xp));
          xv[0] = t.x; xv[KPT - 1] = t.y;
        } else {
          xv[0] = __ldg(xp);
        }
        float2 d2[KPT];
#pragma unroll
        for (int j = 0; j < KPT; ++j) d2[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int n4 = 0; n4 < NT; n4 += 4) {
          const float4 gv = *reinterpret_cast<const float4*>(sg + m * NT + n4);
          const float2 g0 = make_float2(gv.x, gv.y), g1 = make_float2(gv.z, gv.w);
#pragma unroll
          for (int j = 0; j < KPT; ++j) {
            const float2 xx = make_float2(xv[j], xv[j]);
            acc2[j][n4 / 2] = __ffma2_rn(xx, g0, acc2[j][n4 / 2]);
            acc2[j][n4 / 2 + 1] = __ffma2_rn(xx, g1, acc2[j][n4 / 2 + 1]);
            if (DGRAD) {
              d2[j] = __ffma2_rn(g0, wr2[j][n4 / 2], d2[j]);
              d2[j] = __ffma2_rn(g1, wr2[j][n4 / 2 + 1], d2[j]);
            }
          }
        }
        if (DGRAD) {
#pragma unroll
          for (int j = 0; j < KPT; ++j) dv[j] = __fadd_rn(d2[j].x, d2[j].y);
          if (KPT == 2) *reinterpret_cast<float2*>(dxp) = make_float2(dv[0], dv[KPT - 1]);
          else *dxp = dv[0];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < KPT; ++j)
#pragma unroll
    for (int n = 0; n < NT / 2; ++n) {
      red[(rg * KB + kk * KPT + j) * (NT + 1) + 2 * n] = acc2[j][n].x;
      red[(rg * KB + kk * KPT + j) * (NT + 1) + 2 * n + 1] = acc2[j][n].y;
    }
  __syncthreads();
  // dW slab [KB][N] is contiguous in memory: write it coalesced
  for (int i = tid; i < kn; i += blockDim.x) {
    const int r = (int)(((unsigned)i * div_n) >> 16), n = i - r * N;
    float t = red[r * (NT + 1) + n];
#pragma unroll
    for (int g2 = 1; g2 < GR; ++g2) t = __fadd_rn(t, red[(g2 * KB + r) * (NT + 1) + n]);
    dw[(size_t)k0 * N + i] = t;
  }
}
bool dense_head_bwd(const float* x, const float* g, const float* w, float* dw, float* dx, int M, int N, int K, float* ws,
                    size_t ws_floats, cudaStream_t s) {
  if (N > kSkinnyN) return false;
  const int NT = N <= 12 ? 12 : 16;
  // two k columns per thread when X rows keep 8-byte alignment
  const int kpt = (K % 2 == 0 && ((uintptr_t)x & 7) == 0 && (!dx || ((uintptr_t)dx & 7) == 0)) ? 2 : 1;
  const int KB = kBwdK * kpt;
  const int GRr = kBwdGroups;  // (8 row groups = 256-thread CTAs need <= 64 registers to keep the 484 CTAs of cfg2 in one wave: not taken)
  const size_t smem = ((size_t)kBwdRows * NT + KB * (NT + 1) + (size_t)GRr * KB * (NT + 1)) * 4;
  if (smem > 48 * 1024) return false;
  const int kblocks = (K + KB - 1) / KB, thr = kBwdK * GRr;
  // enough CTAs for ~2 waves at 7 CTAs/SM: split the rows (in whole 128-row staging chunks) when K alone does not give them
  int row_gran = kBwdRows;
  int splits = std::max(1, std::min({(2 * 7 * kNumSMs + kblocks - 1) / kblocks, (M + kBwdRows - 1) / kBwdRows,
                                     (int)(ws_floats / ((size_t)K * N))}));
  // still fewer CTAs than SMs (a short K and a small batch, e.g. config 5's 1024 -> 10 head at batch 256: 16 x 2): split the rows
  // finer than the staging chunk, up to about one CTA per SM (config 5: 32 -> 128 CTAs)
  if (kblocks * splits < kNumSMs && M > 32) {
    row_gran = 32;
    splits = std::max(1, std::min({kNumSMs / kblocks, (M + 31) / 32, (int)(ws_floats / ((size_t)K * N))}));
  }
  const int rows_per_split = ((M + splits - 1) / splits + row_gran - 1) / row_gran * row_gran;
  splits = (M + rows_per_split - 1) / rows_per_split;
  float* dwo = splits > 1 ? ws : dw;
  const dim3 grid(kblocks, splits);
  const int pw = prewait_ok();
#define SB_HEAD_BWD(NT_, DG_, KP_) \
  launch_k(dense_head_bwd_kernel<NT_, DG_, KP_>, grid, thr, smem, s, x, g, w, dwo, dx, M, K, N, rows_per_split, pdl_early_hint(), pw)
  // rows of X in flight per thread: the kernel is bound by L2 latency x bytes in flight (13 warps per SM at ~90 registers), so the
  // row loop is unrolled over ALL the rows a thread owns when they are few (cfg2: 100 rows / 4 groups = 25: 1020.6 k -> 1033.6 k
  // samples/s against 8 in flight; 13 in flight: 1022 k)
  const int rows_per_thread = (std::min({M, kBwdRows, rows_per_split}) + kBwdGroups - 1) / kBwdGroups;
  if (NT == 12 && kpt == 2 && dx && rows_per_thread > 16 && rows_per_thread <= 25) {
    launch_k(dense_head_bwd_kernel<12, true, 2, kBwdGroups, 25>, grid, thr, smem, s, x, g, w, dwo, dx, M, K, N, rows_per_split, pdl_early_hint(), pw);
  } else if (NT == 12 && kpt == 2 && dx && rows_per_thread > 25) {
    launch_k(dense_head_bwd_kernel<12, true, 2, kBwdGroups, 32>, grid, thr, smem, s, x, g, w, dwo, dx, M, K, N, rows_per_split, pdl_early_hint(), pw);
  } else if (NT == 12) {
    if (kpt == 2) { if (dx) SB_HEAD_BWD(12, true, 2); else SB_HEAD_BWD(12, false, 2); }
    else { if (dx) SB_HEAD_BWD(12, true, 1); else SB_HEAD_BWD(12, false, 1); }
  } else {
    if (kpt == 2) { if (dx) SB_HEAD_BWD(16, true, 2); else SB_HEAD_BWD(16, false, 2); }
    else { if (dx) SB_HEAD_BWD(16, true, 1); else SB_HEAD_BWD(16, false, 1); }
  }
#undef SB_HEAD_BWD
  SB_LAUNCHED();
  if (splits > 1) partial_reduce(ws, dw, K * N, splits, s);
  return true;
}

// =====================================================================================================================
// Bias gradient fused with the backward of the in-place activation that follows the Bias (layer/Bias.scala:32-33 under
// layer/Activate.scala:16-17):  g <- act'(y) * g  (in place)  and  db[c] = sum_rows g[r][c]  in ONE pass over g and y.
// Block = 32 column quads (128 columns) x 8 row lanes over a chunk of rows; per-chunk partials, summed in chunk order.
// act = 0 gives the plain column sum (g untouched, y unused).
// =====================================================================================================================
__device__ __forceinline__ float act_grad1(float g, float y, int act, float thr) {
  switch (act) {
    case 1: return __fmul_rn(__fmul_rn(y, __fsub_rn(1.f, y)), g);
    case 2: return __fmul_rn(__fsub_rn(1.f, __fmul_rn(y, y)), g);
    case 4: return y > thr ? g : 0.f;
    default: return g;
  }
}
__global__ void __launch_bounds__(256) act_bwd_col_sum_kernel(float* __restrict__ g, const float* __restrict__ y, int act, float thr,
                                                              float* __restrict__ partial, long long rows, int C,
                                                              long long rows_per_chunk) {
  pdl_prologue();
  __shared__ float4 red[8][33];
  const int c4 = blockIdx.x * 32 + threadIdx.x;  // column quad
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 * 4 < C) {
#pragma unroll 4
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      float4 gv = *reinterpret_cast<const float4*>(g + r * C + c4 * 4);
      if (act) {
        const float4 yv = ld4(y + r * C + c4 * 4);
        gv.x = act_grad1(gv.x, yv.x, act, thr); gv.y = act_grad1(gv.y, yv.y, act, thr);
        gv.z = act_grad1(gv.z, yv.z, act, thr); gv.w = act_grad1(gv.w, yv.w, act, thr);
        st4(g + r * C + c4 * 4, gv);
      }
      acc.x = __fadd_rn(acc.x, gv.x); acc.y = __fadd_rn(acc.y, gv.y); acc.z = __fadd_rn(acc.z, gv.z); acc.w = __fadd_rn(acc.w, gv.w);
    }
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c4 * 4 < C) {
    for (int j = 1; j < 8; ++j) {
      const float4 v = red[j][threadIdx.x];
      acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
    }
    st4(partial + (long long)blockIdx.y * C + c4 * 4, acc);
  }
}
// returns false when the shape is not covered (C % 4 != 0): the caller runs act_bwd + col_sum
bool act_bwd_col_sum(float* g, const float* y, int act, float thr, float* out, long long rows, int C, float* ws, size_t ws_floats,
                     cudaStream_t s) {
  if ((C & 3) || rows < 1) return false;
  const int col_blocks = (C / 4 + 31) / 32;
  long long chunks = (4LL * kNumSMs + col_blocks - 1) / col_blocks;
  if (chunks > (rows + 31) / 32) chunks = (rows + 31) / 32;
  if (chunks > (long long)(ws_floats / (size_t)C)) chunks = (long long)(ws_floats / (size_t)C);
  if (chunks < 1) return false;
  const long long rpc = (rows + chunks - 1) / chunks;
  chunks = (rows + rpc - 1) / rpc;
  launch_k(act_bwd_col_sum_kernel, dim3(col_blocks, (unsigned)chunks), dim3(32, 8), 0, s, g, y, act, thr, ws, rows, C, rpc);
  SB_LAUNCHED();
  partial_reduce(ws, out, C, (int)chunks, s);
  return true;
}

// Fused head: z = (sum of the split-K partials | z) + bias ; p = exp(z)/sum ; loss = -sum(y log(p+eps))/B ;
// dz = -(p/B)(r - sum r p), r = y/(p+eps) ; db = column sums of dz.  CTA g owns rows [g*R, g*R+R): 8 thread groups sum the
// partials of every (row, class) (chunk c goes to group c % 8, then the groups in order), one warp per row does the softmax /
// loss / dz; per-CTA loss and bias-gradient partials go to `scratch` and the LAST CTA to finish (ticket) adds them in CTA order,
// so the result does not depend on which CTA that is.  (layer/Bias.scala:32-33, models/Activation.scala:63-69,
// models/Loss.scala:44-57)
constexpr int kHeadMaxRows = 32;   // rows per CTA (one warp per row)
constexpr int kHeadMaxCnt = 128;   // (row, class) outputs per CTA
__global__ void __launch_bounds__(1024) head_bias_softmax_cce_kernel(float* __restrict__ z, const float* __restrict__ partial,
                                                                     int chunks, const float* __restrict__ bias,
                                                                     const float* __restrict__ y, float* __restrict__ dz,
                                                                     float* __restrict__ dbias, StepState* st, int rows, int C,
                                                                     int rows_per_cta, float* __restrict__ scratch,
                                                                     unsigned int* ticket, int early, int defer_combine) {
  pdl_prologue_trigger_if(early);
  __shared__ float zs[32][kHeadMaxCnt];
  __shared__ float sl[kHeadMaxRows];
  __shared__ float sdb[kHeadMaxRows][33];
  __shared__ bool last;
  const float eps = 1e-8f, fr = (float)rows;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int row0 = blockIdx.x * rows_per_cta, nrows = min(rows_per_cta, rows - row0);
  const int mn = rows * C, cnt = nrows * C;
  // labels and bias of this warp's row: requested now, consumed after the barrier (one L2 round trip less on the critical path)
  float yv_pre = 0.f, bias_pre = 0.f;
  if (wid < nrows && lane < C) {
    yv_pre = __ldg(y + (size_t)(row0 + wid) * C + lane);
    if (bias) bias_pre = __ldg(bias + lane);
  }
  // ---- z for this CTA's rows: L lanes (one per output) x G groups; chunk c goes to group c % G, groups are added in order ----
  const int L = (cnt + 31) & ~31, G = 1024 / L;
  {
    const int grp = tid / L, i = tid - grp * L;
    if (grp < G && i < cnt) {
      float v = 0.f;
      if (partial) {
        const float* pp = partial + (size_t)row0 * C + i;
        for (int c = grp; c < chunks; c += G) v = __fadd_rn(v, __ldg(pp + (size_t)c * mn));
      } else if (grp == 0) {
        v = z[(size_t)row0 * C + i];
      }
      zs[grp][i] = v;
    }
  }
  __syncthreads();
  float l = 0.f, d = 0.f;
  if (wid < nrows) {
    const int row = row0 + wid;
    float e = 0.f, yv = 0.f;
    if (lane < C) {
      float v = zs[0][wid * C + lane];
      for (int g2 = 1; g2 < G; ++g2) v = __fadd_rn(v, zs[g2][wid * C + lane]);
      if (bias) v = __fadd_rn(v, bias_pre);
      e = expf(v);
      yv = yv_pre;
    }
    float sum = e;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
    const float p = __fdiv_rn(e, sum);
    const float pe = __fadd_rn(p, eps);
    l = lane < C ? __fmul_rn(yv, logf(pe)) : 0.f;
    const float r = lane < C ? __fdiv_rn(yv, pe) : 0.f;
    float dot = __fmul_rn(r, p);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l = __fadd_rn(l, __shfl_xor_sync(0xffffffffu, l, o));
      dot = __fadd_rn(dot, __shfl_xor_sync(0xffffffffu, dot, o));
    }
    if (lane < C) {
      z[(size_t)row * C + lane] = p;
      if (dz) {
        d = __fmul_rn(-__fdiv_rn(p, fr), __fsub_rn(r, dot));
        dz[(size_t)row * C + lane] = d;
      }
    }
    if (lane == 0) sl[wid] = l;
    sdb[wid][lane] = d;
  }
  __syncthreads();
  // ---- per-CTA partials -> scratch[cta][0] = loss sum, scratch[cta][1 + c] = bias gradient ----
  float* mine = scratch + (size_t)blockIdx.x * 40;
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < nrows; ++i) t = __fadd_rn(t, sl[i]);
    mine[0] = t;
  }
  if (tid >= 32 && tid < 32 + C) {
    float t = 0.f;
    for (int i = 0; i < nrows; ++i) t = __fadd_rn(t, sdb[i][tid - 32]);
    mine[1 + tid - 32] = t;
  }
  // defer_combine: the per-CTA partials are final when this kernel completes; head_loss_combine_kernel adds them on the backward
  // side branch -- dz is all the next kernel of the step needs, and the ticket + fence + reload below kept ~2 us of pure latency
  // on the step's critical path (in-situ timeline, profiles/r2_trace_cfg2.txt: 7.4 us for this kernel)
  if (defer_combine) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the LAST CTA to finish adds the per-CTA partials in CTA order (one warp per quantity, lanes stride the CTAs, fixed tree)
  const int q = wid;  // 0: loss, 1..C: bias gradient of class q-1
  if (q <= C) {
    float t = 0.f;
    for (unsigned g2 = lane; g2 < gridDim.x; g2 += 32) t = __fadd_rn(t, __ldcg(scratch + (size_t)g2 * 40 + q));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (lane == 0) {
      if (q == 0) st->loss = __fdiv_rn(-t, fr);
      else if (dz && dbias) dbias[q - 1] = t;
    }
  }
  if (tid == 0) *ticket = 0;  // ready for the next launch
}
// the same final sum as the last-CTA path above (one warp per quantity, lanes stride the CTAs, fixed tree): same bits
__global__ void __launch_bounds__(1024) head_loss_combine_kernel(const float* __restrict__ scratch, int n_ctas, int C, int rows,
                                                                 StepState* st, float* __restrict__ dbias) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  if (q > C) return;
  float t = 0.f;
  for (int g2 = lane; g2 < n_ctas; g2 += 32) t = __fadd_rn(t, __ldcg(scratch + (size_t)g2 * 40 + q));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, o));
  if (lane == 0) {
    if (q == 0) st->loss = __fdiv_rn(-t, (float)rows);
    else if (dbias) dbias[q - 1] = t;
  }
}
void head_loss_combine(const float* scratch, int n_ctas, int C, int rows, StepState* st, float* dbias, cudaStream_t s) {
  launch_k(head_loss_combine_kernel, 1, 32 * (C + 1), 0, s, scratch, n_ctas, C, rows, st, dbias);
  SB_LAUNCHED();
}
// scratch: 1024 x 40 floats of per-CTA partials followed by one zero-initialised 32-bit ticket (the kernel leaves it zero)
bool head_bias_softmax_cce(float* z, const float* partial, int chunks, const float* bias, const float* y, float* dz, float* dbias,
                           StepState* st, int rows, int C, float* scratch, cudaStream_t s, int* deferred_ctas) {
  if (C > 32 || rows < 1) return false;
  // few rows per CTA (a short dependent-load chain over the split-K partials), but at most 1024 CTAs
  int rpc = std::max(std::max(1, 32 / C), (rows + 1023) / 1024);
  if (rpc > kHeadMaxRows || rpc * C > kHeadMaxCnt) return false;
  const int grid = (rows + rpc - 1) / rpc;
  launch_k(head_bias_softmax_cce_kernel, grid, 1024, 0, s, z, partial, chunks, bias, y, dz, dbias, st, rows, C, rpc, scratch,
           reinterpret_cast<unsigned int*>(scratch + 1024 * 40), pdl_early_hint(), deferred_ctas ? 1 : 0);
  SB_LAUNCHED();
  if (deferred_ctas) *deferred_ctas = grid;
  return true;
}

}  // namespace sb

namespace sb {
void trace_set_kernels_fast(unsigned long long* p) { sb_trace_set_local(p); }  // SB_TRACE timeline buffer of this translation unit
}  // namespace sb
