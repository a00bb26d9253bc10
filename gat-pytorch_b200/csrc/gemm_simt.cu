// fp32 FFMA GEMM (algo 1): exact-fp32 fallback for shapes the tcgen05 path cannot take
// (leading dimensions not 16-byte aligned, e.g. Cora's K=1433, PATTERN's K=3) and the parity
// baseline the 3xTF32 kernel is checked against.  Row-major, any transposition, any size.
// Deterministic: split-K partials are written to the workspace and summed in a fixed order.
#include "common.cuh"

namespace gat {

constexpr int kBK = 16;

template <int BM, int BN, bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(int64_t M, int64_t N, int64_t K, const float* __restrict__ A, int64_t lda,
                 const float* __restrict__ B, int64_t ldb, float* __restrict__ C, int64_t ldc,
                 int64_t k_chunk, int64_t c_split_stride, int act_a, int act_b, const float* __restrict__ mul_src, int64_t mul_ld) {
  constexpr int TM = BM / 16, TN = BN / 16;
  __shared__ __align__(16) float As[kBK][BM + 4];
  __shared__ __align__(16) float Bs[kBK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t k_begin = (int64_t)blockIdx.z * k_chunk;
  const int64_t k_end = min(K, k_begin + k_chunk);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = k_begin; k0 < k_end; k0 += kBK) {
    // A tile: BM x kBK
#pragma unroll
    for (int i = 0; i < BM * kBK / 256; ++i) {
      int idx = tid + i * 256;
      int m, k;
      if (TA) { m = idx % BM; k = idx / BM; } else { k = idx % kBK; m = idx / kBK; }
      int64_t gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < k_end) v = TA ? A[gk * lda + gm] : A[gm * lda + gk];
      if (act_a) v = v > 0.f ? v : expm1f(v);        // fused glue: ELU on the operand
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < BN * kBK / 256; ++i) {
      int idx = tid + i * 256;
      int n, k;
      if (TB) { k = idx % kBK; n = idx / kBK; } else { n = idx % BN; k = idx / BN; }
      int64_t gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < k_end) v = TB ? B[gn * ldb + gk] : B[gk * ldb + gn];
      if (act_b) v = v > 0.f ? v : expm1f(v);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < TM / 4; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&As[k][ty * 4 + g * 64]);
        a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < TN / 4; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[k][tx * 4 + g * 64]);
        b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* Cz = C + (int64_t)blockIdx.z * c_split_stride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t gm = m0 + ty * 4 + (i / 4) * 64 + (i % 4);
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int64_t gn = n0 + tx * 4 + (j / 4) * 64 + (j % 4);
      if (gn < N) {
        float v = acc[i][j];
        if (mul_src != nullptr) { const float x = mul_src[gm * mul_ld + gn]; v *= x > 0.f ? 1.f : expf(x); }   // ELU'(x)
        Cz[gm * ldc + gn] = v;
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ P, int splits, int64_t M, int64_t N,
                                     float* __restrict__ C, int64_t ldc, const float* __restrict__ mul_src, int64_t mul_ld) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  double s = 0.0;
  for (int z = 0; z < splits; ++z) s += (double)P[(int64_t)z * M * N + idx];
  float v = (float)s;
  if (mul_src != nullptr) { const float x = mul_src[(idx / N) * mul_ld + (idx % N)]; v *= x > 0.f ? 1.f : expf(x); }
  C[(idx / N) * ldc + (idx % N)] = v;
}

namespace tc {
void splitk_reduce_launch(const float* partial, int splits, int64_t m, int64_t n, float* c, int64_t ldc, cudaStream_t st) {
  const int64_t total = m * n;
  splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, splits, m, n, c, ldc, nullptr, 0);
}
}  // namespace tc

struct SimtPlan { int bm, bn, splits; int64_t k_chunk; };

SimtPlan simt_plan(int64_t m, int64_t n, int64_t k) {
  SimtPlan p;
  p.bm = m <= 64 ? 64 : 128;
  p.bn = n <= 64 ? 64 : 128;
  int64_t tiles = ((m + p.bm - 1) / p.bm) * ((n + p.bn - 1) / p.bn);
  p.splits = 1;
  // few tiles and a long contraction (dW of the small graphs: 64 x 1433 over K = 2708 nodes; Cora's projection: 22 tiles over
  // K = 1433): split K so the grid covers the SMs; a split never gets fewer than 128 k-steps
  if (tiles < kNumSMs && k >= 512) {
    int64_t want = (2 * kNumSMs + tiles - 1) / tiles;
    int64_t cap = k / 128;
    p.splits = (int)(want < cap ? want : cap);
    if (p.splits < 1) p.splits = 1;
    if (p.splits > 1024) p.splits = 1024;
  }
  int64_t chunk = (k + p.splits - 1) / p.splits;
  chunk = (chunk + kBK - 1) / kBK * kBK;
  if (chunk < kBK) chunk = kBK;
  p.k_chunk = chunk;
  p.splits = (int)((k + chunk - 1) / chunk);
  if (p.splits < 1) p.splits = 1;
  return p;
}

size_t simt_workspace_bytes(int64_t m, int64_t n, int64_t k) {
  SimtPlan p = simt_plan(m, n, k);
  return p.splits > 1 ? (size_t)p.splits * (size_t)m * (size_t)n * sizeof(float) : 0;
}

template <int BM, int BN>
static void launch_simt(bool ta, bool tb, dim3 grid, cudaStream_t st, int64_t M, int64_t N, int64_t K, const float* A,
                        int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t kc, int64_t cs,
                        int aa, int ab, const float* ms, int64_t ml) {
  if (!ta && !tb) gemm_simt_kernel<BM, BN, false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, kc, cs, aa, ab, ms, ml);
  else if (!ta && tb) gemm_simt_kernel<BM, BN, false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, kc, cs, aa, ab, ms, ml);
  else if (ta && !tb) gemm_simt_kernel<BM, BN, true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, kc, cs, aa, ab, ms, ml);
  else gemm_simt_kernel<BM, BN, true, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, kc, cs, aa, ab, ms, ml);
}

int gemm_simt(int ta, int tb, int64_t m, int64_t n, int64_t k, const float* a, int64_t lda, const float* b,
              int64_t ldb, float* c, int64_t ldc, void* workspace, size_t workspace_bytes, cudaStream_t st,
              int act_a, int act_b, const float* mul_src, int64_t mul_ld) {
  if (m == 0 || n == 0) return GAT_OK;
  SimtPlan p = simt_plan(m, n, k);
  size_t need = p.splits > 1 ? (size_t)p.splits * m * n * sizeof(float) : 0;
  if (need > workspace_bytes || (need && !workspace)) {
    set_error("gat_gemm: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GAT_EWORKSPACE;
  }
  dim3 grid((unsigned)((n + p.bn - 1) / p.bn), (unsigned)((m + p.bm - 1) / p.bm), (unsigned)p.splits);
  float* out = p.splits > 1 ? (float*)workspace : c;
  int64_t out_ld = p.splits > 1 ? n : ldc;
  int64_t cs = p.splits > 1 ? m * n : 0;
  const float* ms = p.splits > 1 ? nullptr : mul_src;    // with split-K the multiplier is applied after the reduction
  if (p.bm == 128 && p.bn == 128) launch_simt<128, 128>(ta, tb, grid, st, m, n, k, a, lda, b, ldb, out, out_ld, p.k_chunk, cs, act_a, act_b, ms, mul_ld);
  else if (p.bm == 128 && p.bn == 64) launch_simt<128, 64>(ta, tb, grid, st, m, n, k, a, lda, b, ldb, out, out_ld, p.k_chunk, cs, act_a, act_b, ms, mul_ld);
  else if (p.bm == 64 && p.bn == 128) launch_simt<64, 128>(ta, tb, grid, st, m, n, k, a, lda, b, ldb, out, out_ld, p.k_chunk, cs, act_a, act_b, ms, mul_ld);
  else launch_simt<64, 64>(ta, tb, grid, st, m, n, k, a, lda, b, ldb, out, out_ld, p.k_chunk, cs, act_a, act_b, ms, mul_ld);
  GAT_LAUNCH_CHECK();
  if (p.splits > 1) {
    int64_t total = m * n;
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float*)workspace, p.splits, m, n, c, ldc, mul_src, mul_ld);
    GAT_LAUNCH_CHECK();
  }
  return GAT_OK;
}

}  // namespace gat
