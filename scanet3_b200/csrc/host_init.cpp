// Host-side pieces of weight initialisation: TF-Philox normal samples + Householder QR for the Orthogonal
// initializer (reference: models/Initializer.scala:202-219, math/rand/AllKernels.scala:23-32,
// math/linalg/AllKernels.scala:38-58).  Runs once at epoch 0; not on the hot path.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "tc.h"

namespace sb {

static void philox_block_host(uint64_t seed, uint64_t ctr, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0, c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static float u32_to_float_host(uint32_t x) {
  uint32_t bits = (127u << 23) | (x & 0x7fffffu);
  float f;
  memcpy(&f, &bits, 4);
  return f - 1.0f;
}
static void philox_normal_host(uint64_t seed, int64_t n, float* out) {
  for (int64_t blk = 0; blk * 4 < n; ++blk) {
    uint32_t w[4];
    philox_block_host(seed, (uint64_t)blk, w);
    float v[4];
    for (int j = 0; j < 4; j += 2) {  // BoxMullerFloat
      float u1 = u32_to_float_host(w[j]);
      if (u1 < 1.0e-7f) u1 = 1.0e-7f;
      float v1 = (float)(2.0 * M_PI * (double)u32_to_float_host(w[j + 1]));
      float u2 = sqrtf(-2.0f * logf(u1));
      v[j] = sinf(v1) * u2;
      v[j + 1] = cosf(v1) * u2;
    }
    for (int j = 0; j < 4 && blk * 4 + j < n; ++j) out[blk * 4 + j] = v[j];
  }
}

void orthogonal_host(float* out, int64_t rows, int64_t cols, unsigned long long seed, float gain) {
  // flat = (max(rows,cols), min(rows,cols)); Q of the reduced QR made unique with sign(diag R); transposed when rows < cols
  const int64_t m = rows > cols ? rows : cols, n = rows > cols ? cols : rows;
  std::vector<float> a0((size_t)(m * n));
  philox_normal_host(seed, m * n, a0.data());
  std::vector<double> A(a0.begin(), a0.end());  // row-major m x n; Householder in double, rounded once at the end
  std::vector<double> Q((size_t)(m * n), 0.0);
  std::vector<std::vector<double>> vs;
  std::vector<double> rdiag((size_t)n, 0.0);
  for (int64_t k = 0; k < n; ++k) {
    double norm = 0;
    for (int64_t i = k; i < m; ++i) norm += A[i * n + k] * A[i * n + k];
    norm = sqrt(norm);
    std::vector<double> v((size_t)m, 0.0);
    double alpha = A[k * n + k] > 0 ? -norm : norm;
    for (int64_t i = k; i < m; ++i) v[i] = A[i * n + k];
    v[k] -= alpha;
    double vn = 0;
    for (int64_t i = k; i < m; ++i) vn += v[i] * v[i];
    if (vn > 0) {
      for (int64_t j = k; j < n; ++j) {
        double dot = 0;
        for (int64_t i = k; i < m; ++i) dot += v[i] * A[i * n + j];
        double f = 2.0 * dot / vn;
        for (int64_t i = k; i < m; ++i) A[i * n + j] -= f * v[i];
      }
    }
    rdiag[k] = A[k * n + k];
    vs.push_back(v);
  }
  // Q = H_0 H_1 ... H_{n-1} [I_n; 0]
  for (int64_t j = 0; j < n; ++j) Q[j * n + j] = 1.0;
  for (int64_t k = n - 1; k >= 0; --k) {
    const std::vector<double>& v = vs[k];
    double vn = 0;
    for (int64_t i = k; i < m; ++i) vn += v[i] * v[i];
    if (vn == 0) continue;
    for (int64_t j = 0; j < n; ++j) {
      double dot = 0;
      for (int64_t i = k; i < m; ++i) dot += v[i] * Q[i * n + j];
      double f = 2.0 * dot / vn;
      for (int64_t i = k; i < m; ++i) Q[i * n + j] -= f * v[i];
    }
  }
  for (int64_t j = 0; j < n; ++j) {
    double sg = rdiag[j] > 0 ? 1.0 : (rdiag[j] < 0 ? -1.0 : 0.0);
    for (int64_t i = 0; i < m; ++i) Q[i * n + j] *= sg;
  }
  const bool transpose = rows < cols;
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t c = 0; c < cols; ++c) {
      double q = transpose ? Q[c * n + r] : Q[r * n + c];
      float f = (float)q;
      if (gain != 0.0f) f *= gain;
      out[r * cols + c] = f;
    }
}

}  // namespace sb
