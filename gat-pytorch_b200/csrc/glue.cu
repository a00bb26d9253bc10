// Caller-side glue next to the layer: the attention-norm regulariser of
// GATModel.calc_attention_norm (GATModel.py:189-234) on the structure Kernel 1 already built.
//
//   norm_l = sum_{e,h} | alpha_l[e,h] * deg(dst_e) - 1 | / E'          (GATModel.py:207-224, one layer)
//   d norm_l / d alpha_l[e,h] = sign(alpha*deg - 1) * deg(dst_e) / E'
//
// The reference builds deg per edge with a scatter_add of ones and an index_select (GATModel.py:196-201), multiplies,
// subtracts and takes torch.norm(p=1): five (E', NH) passes plus their autograd.  Here deg(dst) is a rowptr difference
// (GATModel.py:196-201 == rowptr[d+1]-rowptr[d]), the forward is one pass over alpha and the backward
// one pass that writes dL/dalpha directly.  The sum is reduced in two fixed-order stages in fp64 (deterministic).
#include "common.cuh"

namespace gat {

constexpr int kNormBlocks = 592;   // 4 per SM

template <typename T>
__global__ void __launch_bounds__(256)
attn_norm_partial_kernel(const T* __restrict__ dst, const int32_t* __restrict__ rowptr, const float* __restrict__ alpha,
                         int64_t n_edges, int nh, double* __restrict__ partials) {
  __shared__ double sh[256];
  const int64_t per = (n_edges + kNormBlocks - 1) / kNormBlocks;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n_edges, lo + per);
  double t = 0.0;
  for (int64_t e = lo + threadIdx.x; e < hi; e += 256) {
    const int64_t d = (int64_t)dst[e];
    const float deg = (float)(__ldg(rowptr + d + 1) - __ldg(rowptr + d));
    const float* a = alpha + e * nh;
    float s = 0.f;
    for (int h = 0; h < nh; ++h) s += fabsf(a[h] * deg - 1.0f);
    t += (double)s;
  }
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(1024)
attn_norm_finalize_kernel(const double* __restrict__ partials, double inv_edges, float* __restrict__ out) {
  __shared__ double sh[1024];
  double t = 0.0;
  for (int i = threadIdx.x; i < kNormBlocks; i += 1024) t += partials[i];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)(sh[0] * inv_edges);
}

template <typename T>
__global__ void attn_norm_bwd_kernel(const T* __restrict__ dst, const int32_t* __restrict__ rowptr, const float* __restrict__ alpha,
                                     int64_t n_edges, int nh, const float* __restrict__ upstream, float inv_edges,
                                     float* __restrict__ grad_alpha) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n_edges * nh) return;
  const int64_t e = idx / nh;
  const int64_t d = (int64_t)dst[e];
  const float deg = (float)(__ldg(rowptr + d + 1) - __ldg(rowptr + d));
  const float v = alpha[idx] * deg - 1.0f;
  const float sgn = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);       // torch's sign(): 0 at 0
  grad_alpha[idx] = __ldg(upstream) * sgn * deg * inv_edges;
}

// ---- the regulariser straight from the layer's score terms (fused form) -----------------------------------
// alpha[e,h] = exp(LeakyReLU(s_src[src_e,h] + s_tgt[d,h] - M)) / (Z[d,h] + eps) is recomputed per CSR slot from what the layer
// keeps anyway (16-byte gathers of s_src, L2 resident), so the (E', NH) attention tensor is never written or read:
//   norm      = sum_{d,e,h} |alpha*deg(d) - 1| / E'                      (value, two fixed-order fp64 stages)
//   tsum[d,h] = sum_{e in row d} alpha[e,h] * sign(alpha[e,h]*deg(d) - 1)
// tsum is all the backward needs besides a scalar: with dL/dalpha[e,h] = c*deg*sign(.) the per-target sum of the softmax
// adjoint, S[d,h] = sum_e alpha*dL/dalpha, gains c*deg(d)*tsum[d,h] (gat_edge_bwd_rowdot_glue) and the per-edge term is formed
// inside the one-pass source-major backward (gat_edge_bwd_fused_norm) -- the regularised training step keeps the ONE-pass
// backward and no per-edge record.
__global__ void __launch_bounds__(256)
attn_norm_scores_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n, const float* __restrict__ s_src,
                        const float* __restrict__ s_tgt, const float* __restrict__ gmax, const float* __restrict__ z, int nh,
                        int const_attention, float* __restrict__ tsum, double* __restrict__ partials) {
  __shared__ double sh[8];
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  const float m = const_attention ? 0.f : __ldg(gmax);
  double total = 0.0;
  for (int64_t row = warp; row < n; row += nwarps) {
    const int start = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float deg = (float)(end - start);
    float st[8], zz[8], ns[8], ts[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      ns[h] = 0.f; ts[h] = 0.f;
      st[h] = (!const_attention && h < nh) ? __ldg(s_tgt + row * nh + h) : 0.f;
      zz[h] = h < nh ? __ldg(z + row * nh + h) + kSoftmaxEps : 1.f;
    }
    for (int j = start + lane; j < end; j += 32) {
      const float* sp = s_src + (int64_t)__ldg(col + j) * nh;
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        if (h < nh) {
          float p = 1.f;
          if (!const_attention) {
            const float t = __ldg(sp + h) + st[h] - m;
            p = expf(t >= 0.f ? t : t * kLeakySlope);
          }
          const float a = p / zz[h];
          const float u = a * deg - 1.0f;
          ns[h] += fabsf(u);
          ts[h] += u > 0.f ? a : (u < 0.f ? -a : 0.f);
        }
      }
    }
    float rowsum = 0.f;
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      if (h < nh) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ns[h] += __shfl_xor_sync(0xffffffffu, ns[h], o);
          ts[h] += __shfl_xor_sync(0xffffffffu, ts[h], o);
        }
        rowsum += ns[h];
        if (lane == 0) tsum[row * nh + h] = ts[h];
      }
    }
    total += (double)rowsum;
  }
  if (lane == 0) sh[threadIdx.x >> 5] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partials[blockIdx.x] = t;
  }
}

// ---- micro-averaged F1 of a multilabel prediction (ppi_gat.py:38 / :48 / :56: sklearn f1_score(y_pred = out > 0, y_true = y,
// average="micro") after a D2H copy of both matrices, 29 of every 36 ms of a PPI training step) as three integer counts on the
// device: TP = #(pred & true), FP = #(pred & !true), FN = #(!pred & true).  Integer sums: exact and order independent.
__global__ void __launch_bounds__(256)
micro_f1_counts_kernel(const float* __restrict__ logits, const float* __restrict__ y, int64_t count, unsigned long long* __restrict__ counts) {
  unsigned int tp = 0, fp = 0, fn = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const bool pred = logits[i] > 0.f, truth = y[i] != 0.f;
    tp += pred && truth; fp += pred && !truth; fn += !pred && truth;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tp += __shfl_xor_sync(0xffffffffu, tp, o); fp += __shfl_xor_sync(0xffffffffu, fp, o); fn += __shfl_xor_sync(0xffffffffu, fn, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (tp) atomicAdd(counts + 0, (unsigned long long)tp);
    if (fp) atomicAdd(counts + 1, (unsigned long long)fp);
    if (fn) atomicAdd(counts + 2, (unsigned long long)fn);
  }
}

// ---- visualisation feed -----------------------------------------------------------------------------
// visualisation/entropy_histograms.py:103-115 loops over every node, masks the whole edge list with
// `target_nodes == node_id` (O(N * E')) and calls scipy.stats.entropy(weights, base=2) on the node's incoming attention;
// visualisation/weight_histograms.py:74-87 does the same masking and multiplies the weights by the neighbourhood size.
// On the CSR of Kernel 1 a node's incoming edges are one contiguous segment (in the reference's edge order), so both are
// one pass: a warp per (node) row, lanes over the row's slots, heads in registers.
//   entropy[i,h]  = -sum_e p log2 p,  p = alpha[e,h] / sum_e alpha[e,h]   (scipy normalises pk; terms with p == 0 are 0)
//   uniform[i]    = log2(deg_i)                                            (entropy of ones(deg)/deg, :115)
//   scaled[j,h]   = alpha[eid[j],h] * deg(row of j)                        (CSR order == the reference's node-major order)
// Rows without incoming edges get entropy 0 and uniform 0 (the reference would divide 0/0 there; it never visits such a
// node because every node has a self-loop after the rewrite).
__global__ void __launch_bounds__(256)
attn_entropy_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid, int64_t n, const float* __restrict__ alpha,
                    int nh, float* __restrict__ entropy, float* __restrict__ uniform) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  for (int64_t row = warp; row < n; row += nwarps) {
    const int start = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    float sum[8], ent[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) { sum[h] = 0.f; ent[h] = 0.f; }
    for (int j = start + lane; j < end; j += 32) {
      const float* a = alpha + (int64_t)__ldg(eid + j) * nh;
#pragma unroll
      for (int h = 0; h < 8; ++h)
        if (h < nh) sum[h] += __ldg(a + h);
    }
#pragma unroll
    for (int h = 0; h < 8; ++h)
      if (h < nh) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], o);
      }
    for (int j = start + lane; j < end; j += 32) {
      const float* a = alpha + (int64_t)__ldg(eid + j) * nh;
#pragma unroll
      for (int h = 0; h < 8; ++h)
        if (h < nh) {
          const float p = __ldg(a + h) / sum[h];
          if (p > 0.f) ent[h] -= p * log2f(p);
        }
    }
#pragma unroll
    for (int h = 0; h < 8; ++h)
      if (h < nh) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ent[h] += __shfl_xor_sync(0xffffffffu, ent[h], o);
      }
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < 8; ++h)
        if (h < nh) entropy[row * nh + h] = end > start ? ent[h] : 0.f;
      uniform[row] = end > start ? log2f((float)(end - start)) : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256)
attn_degree_scaled_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid, int64_t n, const float* __restrict__ alpha,
                          int nh, float* __restrict__ scaled) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  for (int64_t row = warp; row < n; row += nwarps) {
    const int start = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float deg = (float)(end - start);
    for (int j = start + lane; j < end; j += 32) {
      const float* a = alpha + (int64_t)__ldg(eid + j) * nh;
      for (int h = 0; h < nh; ++h) scaled[(int64_t)j * nh + h] = __ldg(a + h) * deg;
    }
  }
}

}  // namespace gat

extern "C" int gat_attention_entropy(const int32_t* rowptr, const int32_t* eid, int64_t n, const float* alpha, int nh,
                                     float* entropy, float* uniform, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(rowptr && eid && alpha && entropy && uniform && n >= 0 && nh >= 1 && nh <= 8, "gat_attention_entropy: bad arguments");
  if (n == 0) return GAT_OK;
  const int64_t want = (n + 7) / 8;
  attn_entropy_kernel<<<(unsigned)(want < kNumSMs * 8 ? want : kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(rowptr, eid, n, alpha, nh, entropy, uniform);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_attention_degree_scaled(const int32_t* rowptr, const int32_t* eid, int64_t n, const float* alpha, int nh,
                                           float* scaled, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(rowptr && eid && alpha && scaled && n >= 0 && nh >= 1, "gat_attention_degree_scaled: bad arguments");
  if (n == 0) return GAT_OK;
  const int64_t want = (n + 7) / 8;
  attn_degree_scaled_kernel<<<(unsigned)(want < kNumSMs * 8 ? want : kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(rowptr, eid, n, alpha, nh, scaled);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

// ---- bf16 variant: bfloat16 copy of a feature matrix for the per-edge gathers (round to nearest even) ----
namespace gat {
__device__ __forceinline__ unsigned bf16_bits(float v) {
  unsigned u = __float_as_uint(v);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (u >> 16) | 0x40u;     // NaN stays NaN
  return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
__global__ void __launch_bounds__(256)
f32_to_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, int64_t count4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(src + i);
    dst[i] = make_uint2(bf16_bits(v.x) | (bf16_bits(v.y) << 16), bf16_bits(v.z) | (bf16_bits(v.w) << 16));
  }
}
// the same rounding, kept in fp32 storage: dst = float(bfloat16(src))
__global__ void __launch_bounds__(256)
f32_round_bf16_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int64_t count4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(src + i);
    dst[i] = make_float4(__uint_as_float(bf16_bits(v.x) << 16), __uint_as_float(bf16_bits(v.y) << 16),
                         __uint_as_float(bf16_bits(v.z) << 16), __uint_as_float(bf16_bits(v.w) << 16));
  }
}
}  // namespace gat

extern "C" int gat_edge_bf16_native(int nh, int fp, int go_shared) {
  const int chunks = nh * fp / 4;
  return (nh >= 1 && nh <= 4 && fp > 0 && fp % 4 == 0 && chunks > 32 && chunks <= 64 && !go_shared) ? 1 : 0;
}

extern "C" int gat_f32_round_bf16(const float* src, float* dst, int64_t count, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(src && dst && count >= 0 && count % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0,
                "gat_f32_round_bf16: count must be a multiple of 4 and the buffers 16-byte aligned");
  if (count == 0) return GAT_OK;
  const int64_t c4 = count / 4, want = (c4 + 255) / 256;
  f32_round_bf16_kernel<<<(unsigned)(want < kNumSMs * 16 ? want : kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)src, (float4*)dst, c4);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_f32_to_bf16(const float* src, void* dst, int64_t count, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(src && dst && count >= 0 && count % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0,
                "gat_f32_to_bf16: count must be a multiple of 4 and the buffers 16 / 8 byte aligned");
  if (count == 0) return GAT_OK;
  const int64_t c4 = count / 4, want = (c4 + 255) / 256;
  f32_to_bf16_kernel<<<(unsigned)(want < kNumSMs * 16 ? want : kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)src, (uint2*)dst, c4);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" size_t gat_attention_norm_workspace_bytes(void) { return (size_t)gat::kNormBlocks * sizeof(double); }

extern "C" int gat_attention_norm_fwd(const void* edge_dst, int index_is_int64, const int32_t* rowptr, const float* alpha,
                                      int64_t n_edges, int nh, float* norm_out, void* workspace, size_t workspace_bytes,
                                      gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(edge_dst && rowptr && alpha && norm_out && n_edges >= 1 && nh >= 1, "gat_attention_norm_fwd: bad arguments");
  GAT_CHECK_ARG(workspace && workspace_bytes >= gat_attention_norm_workspace_bytes(), "gat_attention_norm_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* partials = (double*)workspace;
  if (index_is_int64) attn_norm_partial_kernel<int64_t><<<kNormBlocks, 256, 0, st>>>((const int64_t*)edge_dst, rowptr, alpha, n_edges, nh, partials);
  else attn_norm_partial_kernel<int32_t><<<kNormBlocks, 256, 0, st>>>((const int32_t*)edge_dst, rowptr, alpha, n_edges, nh, partials);
  GAT_LAUNCH_CHECK();
  attn_norm_finalize_kernel<<<1, 1024, 0, st>>>(partials, 1.0 / (double)n_edges, norm_out);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_micro_f1_counts(const float* logits, const float* y_true, int64_t count, unsigned long long* counts, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(logits && y_true && counts && count >= 0, "gat_micro_f1_counts: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  GAT_CUDA(cudaMemsetAsync(counts, 0, 3 * sizeof(unsigned long long), st));
  if (count == 0) return GAT_OK;
  const int64_t want = (count + 255) / 256;
  micro_f1_counts_kernel<<<(unsigned)(want < kNumSMs * 8 ? want : kNumSMs * 8), 256, 0, st>>>(logits, y_true, count, counts);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_attention_norm_scores(const int32_t* rowptr, const int32_t* col, int64_t n, int64_t n_edges, const float* s_src,
                                         const float* s_tgt, const float* gmax, const float* z, int nh, int const_attention,
                                         float* tsum, float* norm_out, void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(rowptr && col && z && tsum && norm_out && n >= 1 && n_edges >= 1 && nh >= 1 && nh <= 8, "gat_attention_norm_scores: bad arguments");
  GAT_CHECK_ARG(const_attention || (s_src && s_tgt && gmax), "gat_attention_norm_scores: score buffers missing");
  GAT_CHECK_ARG(workspace && workspace_bytes >= gat_attention_norm_workspace_bytes(), "gat_attention_norm_scores: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  attn_norm_scores_kernel<<<kNormBlocks, 256, 0, st>>>(rowptr, col, n, s_src, s_tgt, gmax, z, nh, const_attention ? 1 : 0, tsum,
                                                       (double*)workspace);
  GAT_LAUNCH_CHECK();
  attn_norm_finalize_kernel<<<1, 1024, 0, st>>>((const double*)workspace, 1.0 / (double)n_edges, norm_out);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_attention_norm_bwd(const void* edge_dst, int index_is_int64, const int32_t* rowptr, const float* alpha,
                                      int64_t n_edges, int nh, const float* upstream, float* grad_alpha, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(edge_dst && rowptr && alpha && upstream && grad_alpha && n_edges >= 1 && nh >= 1, "gat_attention_norm_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = n_edges * nh;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  const float inv = (float)(1.0 / (double)n_edges);
  if (index_is_int64) attn_norm_bwd_kernel<int64_t><<<blocks, 256, 0, st>>>((const int64_t*)edge_dst, rowptr, alpha, n_edges, nh, upstream, inv, grad_alpha);
  else attn_norm_bwd_kernel<int32_t><<<blocks, 256, 0, st>>>((const int32_t*)edge_dst, rowptr, alpha, n_edges, nh, upstream, inv, grad_alpha);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Parameter packing (one launch per direction instead of a dozen tiny torch view / pad / slice / cat kernels).
//   W   (NH*F, F_in)   -> W_p  (NH*Fp, F_in)  padded-head rows (only when Fp != F) and W_pT (F_in, NH*Fp), the K-major
//                         operand of the dX product, so no transpose kernel runs in the backward
//   a   (NH, NH*2F)    -> A_src_p, A_tgt_p (NH, NH*Fp): a.view(NH, NH, 2F)[:, :, :F] / [:, :, F:] (gat_layer.py:76-82:
//                         column h'*2F + j multiplies Wh[src, h', j], column h'*2F + F + j multiplies Wh[dst, h', j])
// and the adjoint (gradients back to the reference's parameter layouts).
// ---------------------------------------------------------------------------------------------------------------
namespace gat {

__global__ void __launch_bounds__(256)
pack_params_kernel(const float* __restrict__ W, const float* __restrict__ a, int nh, int f, int fp, int64_t f_in,
                   float* __restrict__ W_p, float* __restrict__ W_pT, float* __restrict__ a_src_p, float* __restrict__ a_tgt_p) {
  const int64_t dp = (int64_t)nh * fp, n_w = dp * f_in, n_a = (int64_t)nh * dp;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (W_p != nullptr) {       // index over W_p (row r = h*fp + j, column k)
    for (int64_t i = t0; i < n_w; i += stride) {
      const int64_t r = i / f_in, k = i - r * f_in;
      const int h = (int)(r / fp), j = (int)(r - (int64_t)h * fp);
      W_p[i] = j < f ? __ldg(W + ((int64_t)h * f + j) * f_in + k) : 0.f;
    }
  }
  if (W_pT != nullptr) {      // index over W_pT (row k, column r): coalesced writes, strided (L2-resident) reads
    for (int64_t i = t0; i < n_w; i += stride) {
      const int64_t k = i / dp, r = i - k * dp;
      const int h = (int)(r / fp), j = (int)(r - (int64_t)h * fp);
      W_pT[i] = j < f ? __ldg(W + ((int64_t)h * f + j) * f_in + k) : 0.f;
    }
  }
  if (a != nullptr) {
    for (int64_t i = t0; i < n_a; i += stride) {
      const int64_t row = i / dp, c = i - row * dp;
      const int h2 = (int)(c / fp), j = (int)(c - (int64_t)h2 * fp);
      const float* base = a + row * ((int64_t)nh * 2 * f) + (int64_t)h2 * 2 * f + j;
      a_src_p[i] = j < f ? __ldg(base) : 0.f;
      a_tgt_p[i] = j < f ? __ldg(base + f) : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256)
unpack_param_grads_kernel(const float* __restrict__ gW_p, const float* __restrict__ ga_src_p, const float* __restrict__ ga_tgt_p,
                          int nh, int f, int fp, int64_t f_in, float* __restrict__ gW, float* __restrict__ ga) {
  const int64_t dp = (int64_t)nh * fp;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gW != nullptr) {
    const int64_t n_w = (int64_t)nh * f * f_in;
    for (int64_t i = t0; i < n_w; i += stride) {
      const int64_t r = i / f_in, k = i - r * f_in;
      const int h = (int)(r / f), j = (int)(r - (int64_t)h * f);
      gW[i] = __ldg(gW_p + ((int64_t)h * fp + j) * f_in + k);
    }
  }
  if (ga != nullptr) {
    const int64_t n_a = (int64_t)nh * nh * 2 * f;
    for (int64_t i = t0; i < n_a; i += stride) {
      const int64_t row = i / ((int64_t)nh * 2 * f), c = i - row * ((int64_t)nh * 2 * f);
      const int h2 = (int)(c / (2 * f)), jj = (int)(c - (int64_t)h2 * 2 * f);
      const int64_t src = row * dp + (int64_t)h2 * fp + (jj < f ? jj : jj - f);
      ga[i] = jj < f ? __ldg(ga_src_p + src) : __ldg(ga_tgt_p + src);
    }
  }
}

// Neighbourhood plot feed (visualisation/neighbourhood_attention_weights.py:45-58): for each requested target node, its
// neighbours' ids (edge-list order) and the attention of one head over them, divided by its maximum and multiplied by
// 60 / neighbourhood size -- the edge widths of the star plot.  One CTA per requested node over its CSR segment; the
// reference builds each with a full-edge-list mask.
__global__ void __launch_bounds__(128)
neighbourhood_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ eid,
                     const float* __restrict__ alpha, int nh, int head, const int64_t* __restrict__ nodes,
                     const int64_t* __restrict__ out_off, int64_t* __restrict__ out_src, float* __restrict__ out_w) {
  const int64_t node = nodes[blockIdx.x];
  const int b = rowptr[node], e = rowptr[node + 1];
  const int64_t o = out_off[blockIdx.x];
  __shared__ float red[4];
  float mx = -INFINITY;
  for (int j = b + threadIdx.x; j < e; j += blockDim.x) mx = fmaxf(mx, __ldg(alpha + (int64_t)eid[j] * nh + head));
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  const float size = (float)(e - b);
  for (int j = b + threadIdx.x; j < e; j += blockDim.x) {
    out_src[o + (j - b)] = col[j];
    out_w[o + (j - b)] = __ldg(alpha + (int64_t)eid[j] * nh + head) / mx * (60.0f / size);   // :58 then :60, same operation order
  }
}

static unsigned pack_grid(int64_t elems) {
  int64_t b = (elems + 255) / 256;
  if (b < 1) b = 1;
  if (b > kNumSMs * 8) b = kNumSMs * 8;
  return (unsigned)b;
}
}  // namespace gat

extern "C" int gat_pack_params(const float* W, const float* a, int nh, int f, int fp, int64_t f_in,
                               float* W_p, float* W_pT, float* a_src_p, float* a_tgt_p, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(W && nh >= 1 && f >= 1 && fp >= f && fp % 4 == 0 && f_in >= 1, "gat_pack_params: bad arguments");
  GAT_CHECK_ARG(a == nullptr || (a_src_p && a_tgt_p), "gat_pack_params: a given without a_src_p / a_tgt_p");
  if (!W_p && !W_pT && !a) return GAT_OK;
  const int64_t dp = (int64_t)nh * fp;
  const int64_t elems = (W_p || W_pT) ? dp * f_in : (int64_t)nh * dp;
  pack_params_kernel<<<pack_grid(elems), 256, 0, (cudaStream_t)stream>>>(W, a, nh, f, fp, f_in, W_p, W_pT, a_src_p, a_tgt_p);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_unpack_param_grads(const float* gW_p, const float* ga_src_p, const float* ga_tgt_p, int nh, int f, int fp,
                                      int64_t f_in, float* gW, float* ga, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(nh >= 1 && f >= 1 && fp >= f && f_in >= 1, "gat_unpack_param_grads: bad arguments");
  GAT_CHECK_ARG((gW == nullptr || gW_p) && (ga == nullptr || (ga_src_p && ga_tgt_p)), "gat_unpack_param_grads: missing packed gradient");
  if (!gW && !ga) return GAT_OK;
  const int64_t elems = gW ? (int64_t)nh * f * f_in : (int64_t)nh * nh * 2 * f;
  unpack_param_grads_kernel<<<pack_grid(elems), 256, 0, (cudaStream_t)stream>>>(gW_p, ga_src_p, ga_tgt_p, nh, f, fp, f_in, gW, ga);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_attention_neighbourhood(const int32_t* rowptr, const int32_t* col, const int32_t* eid, const float* alpha, int nh,
                                           int head, const int64_t* nodes, int64_t n_nodes_req, const int64_t* out_off,
                                           int64_t* out_src, float* out_w, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(rowptr && col && eid && alpha && nodes && out_off && out_src && out_w, "gat_attention_neighbourhood: null buffer");
  GAT_CHECK_ARG(nh >= 1 && head >= 0 && head < nh && n_nodes_req >= 0, "gat_attention_neighbourhood: bad head / count");
  if (n_nodes_req == 0) return GAT_OK;
  neighbourhood_kernel<<<(unsigned)n_nodes_req, 128, 0, (cudaStream_t)stream>>>(rowptr, col, eid, alpha, nh, head, nodes, out_off, out_src, out_w);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}
