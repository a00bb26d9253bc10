// Per-node attention score terms (Kernel 2 epilogue, standalone form).
//
// The reference multiplies the (E', NH*2F) pair tensor by the full cross-head matrix `a`
// (gat_layer.py:76-82).  That product decomposes exactly into per-node terms
//   s_src = Wh * A_src^T,  s_tgt = Wh * A_tgt^T   (n, NH),   logit[e] = s_src[src_e] + s_tgt[dst_e]
// with A_src/A_tgt the (NH, NH*F) halves of `a`.  The dot products are
// accumulated in fp64 from the fp32 Wh: with the committed PATTERN weights the logits reach
// +-1.9e3 and sit at the fp32 noise floor of the 1e-5 parity bar, so this
// tiny contraction (2*NH columns) is kept more accurate than fp32.
#include "common.cuh"

namespace gat {

constexpr int kScoreKC = 128;  // columns of Wh staged per step
constexpr int kScoreRows = 32; // rows per CTA

__global__ void __launch_bounds__(256)
scores_kernel(const float* __restrict__ wh, int64_t n, int dp, const float* __restrict__ a_src,
              const float* __restrict__ a_tgt, int nh, float* __restrict__ s_src, float* __restrict__ s_tgt) {
  __shared__ float wh_s[kScoreRows][kScoreKC + 1];
  __shared__ float a_s[16][kScoreKC + 1];  // rows 0..nh-1 = A_src, nh..2nh-1 = A_tgt
  const int tid = threadIdx.x;
  const int nj = 2 * nh;                      // <= 16
  const int64_t row0 = (int64_t)blockIdx.x * kScoreRows;
  // thread -> (row r, output column j); 256 threads cover 32 rows x 8 columns per pass
  const int r = tid >> 3, jl = tid & 7;
  double acc0 = 0.0, acc1 = 0.0;              // columns jl and jl+8
  for (int k0 = 0; k0 < dp; k0 += kScoreKC) {
    int kc = min(kScoreKC, dp - k0);
    for (int idx = tid; idx < kScoreRows * kScoreKC; idx += 256) {
      int rr = idx / kScoreKC, kk = idx % kScoreKC;
      int64_t gr = row0 + rr;
      wh_s[rr][kk] = (gr < n && kk < kc) ? wh[gr * dp + k0 + kk] : 0.f;
    }
    for (int idx = tid; idx < nj * kScoreKC; idx += 256) {
      int j = idx / kScoreKC, kk = idx % kScoreKC;
      float v = 0.f;
      if (kk < kc) v = (j < nh) ? a_src[(int64_t)j * dp + k0 + kk] : a_tgt[(int64_t)(j - nh) * dp + k0 + kk];
      a_s[j][kk] = v;
    }
    __syncthreads();
    if (jl < nj) {
#pragma unroll 4
      for (int kk = 0; kk < kc; ++kk) acc0 = fma((double)wh_s[r][kk], (double)a_s[jl][kk], acc0);
    }
    if (jl + 8 < nj) {
#pragma unroll 4
      for (int kk = 0; kk < kc; ++kk) acc1 = fma((double)wh_s[r][kk], (double)a_s[jl + 8][kk], acc1);
    }
    __syncthreads();
  }
  int64_t gr = row0 + r;
  if (gr < n) {
    if (jl < nj) {
      if (jl < nh) s_src[gr * nh + jl] = (float)acc0; else s_tgt[gr * nh + (jl - nh)] = (float)acc0;
    }
    int j1 = jl + 8;
    if (j1 < nj) {
      if (j1 < nh) s_src[gr * nh + j1] = (float)acc1; else s_tgt[gr * nh + (j1 - nh)] = (float)acc1;
    }
  }
}

}  // namespace gat

extern "C" int gat_scores_fwd(const float* wh, int64_t n, int dp, const float* a_src, const float* a_tgt, int nh,
                              float* s_src, float* s_tgt, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(nh >= 1 && nh <= 8, "gat_scores_fwd: num_heads %d not in [1, 8]", nh);
  GAT_CHECK_ARG(dp > 0 && n >= 0, "gat_scores_fwd: bad shape");
  if (n == 0) return GAT_OK;
  unsigned blocks = (unsigned)((n + kScoreRows - 1) / kScoreRows);
  scores_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(wh, n, dp, a_src, a_tgt, nh, s_src, s_tgt);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}
