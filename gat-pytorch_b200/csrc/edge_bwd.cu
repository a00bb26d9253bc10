// Kernel 4: atomic-free deterministic backward of the edge stage (autograd of gat_layer.py:70-132;
// formulas in SURVEY.md section 9.2).  Two passes:
//   dst pass over the target-sorted CSR  -> per-edge records {g, m*alpha}, ds_tgt, partial sums of g
//   src pass over the source-sorted CSR  -> d_wh (total), ds_src, arg-max correction
// Every sum is a fixed-order register / shuffle / shared-memory reduction; no floating-point atomics.
#include "edge_common.cuh"

namespace gat {

struct BwdHeader {          // first 256 bytes of the backward workspace
  double gamma;             // sum over all (e,h) of g
  float corr;               // gamma / |T|   (0 when the arg-max set is empty)
  int n_partials;           // number of per-block partials written by the dst pass
};
constexpr size_t kBwdHeaderBytes = 256;

struct EdgeBwdDstParams {
  const int32_t* rowptr; const int32_t* col; const int32_t* eid; int64_t n;
  const float* wh; int nh; int dp; int chunks; int chunks_per_head;
  const float* s_src; const float* s_tgt; const float* gmax; const float* z;
  int const_attention; float dropout_p; uint64_t seed; uint64_t offset;
  const float* go; const float* grad_alpha;
  float* rec; float* ds_tgt; BwdHeader* header; double* partials;
};

template <int G, int SLOTS>
__global__ void __launch_bounds__(kEdgeThreads)
edge_bwd_dst_kernel(const EdgeBwdDstParams P) {
  constexpr int TB = (G < 8) ? G : (SLOTS >= 6 ? 4 : 8);     // edges per transpose-reduce sub-batch
  constexpr int U = (SLOTS >= 4) ? 2 : (TB < 4 ? TB : 4);    // edges in flight
  constexpr int GROUPS = kEdgeThreads / G;
  extern __shared__ float dyn_smem[];
  __shared__ int sh_src[kEdgeThreads];
  __shared__ float sh_da[kEdgeThreads * kMaxHeads];
  __shared__ double sh_gamma[kEdgeThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, gl = tid & (G - 1), gbase = tid - gl;
  const unsigned gmask = group_mask<G>(lane);
  const int64_t row = (int64_t)blockIdx.x * GROUPS + tid / G;
  const int nh = P.nh;
  const int pstride = P.chunks + 1;
  float* part = dyn_smem + (size_t)(tid / G) * TB * pstride;   // [TB][chunks+1] of my group
  double gam = 0.0;

  if (row < P.n) {
    bool ok[SLOTS];
    float4 go[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      int c = s * G + gl;
      ok[s] = c < P.chunks;
      go[s] = ok[s] ? ldg4(P.go + row * P.dp + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int start = __ldg(P.rowptr + row), end = __ldg(P.rowptr + row + 1);
    float st[kMaxHeads], z[kMaxHeads];
    float gmax = 0.f;
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) {
      st[h] = 0.f;
      z[h] = h < nh ? __ldg(P.z + row * nh + h) : 0.f;
    }
    if (!P.const_attention) {
      gmax = __ldg(P.gmax);
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h) st[h] = h < nh ? __ldg(P.s_tgt + row * nh + h) : 0.f;
    }
    const bool single = (end - start) <= G;
    float alpha[kMaxHeads], dal[kMaxHeads], msk[kMaxHeads], ssum[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) { alpha[h] = 0.f; dal[h] = 0.f; msk[h] = 1.f; ssum[h] = 0.f; }

    // ---- pass 1: d_alpha[e,h] = m * <go[i,h,:], wh[src,h,:]> + grad_alpha;  S[h] = sum alpha*d_alpha
    for (int base = start; base < end; base += G) {
      const int e = base + gl;
      const bool valid = e < end;
      int my_src = 0;
      float ga[kMaxHeads];
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h) { alpha[h] = 0.f; msk[h] = 1.f; ga[h] = 0.f; }
      if (valid) {
        my_src = __ldg(P.col + e);
        if (P.const_attention) {
#pragma unroll
          for (int h = 0; h < kMaxHeads; ++h) alpha[h] = h < nh ? 1.f / (z[h] + kSoftmaxEps) : 0.f;
        } else {
          const float* ss = P.s_src + (int64_t)my_src * nh;
#pragma unroll
          for (int h = 0; h < kMaxHeads; ++h)
            if (h < nh) alpha[h] = attn_exp(__ldg(ss + h) + st[h], gmax) / (z[h] + kSoftmaxEps);
        }
        if (P.dropout_p > 0.f || P.grad_alpha) {
          const int edge_id = __ldg(P.eid + e);
          if (P.dropout_p > 0.f) dropout_scales(P.seed, P.offset, (uint32_t)edge_id, nh, P.dropout_p, msk);
          if (P.grad_alpha) {
#pragma unroll
            for (int h = 0; h < kMaxHeads; ++h)
              if (h < nh) ga[h] = __ldg(P.grad_alpha + (int64_t)edge_id * nh + h);
          }
        }
        sh_src[tid] = my_src;
      }
      __syncwarp(gmask);
      const int cnt = min(G, end - base);
      for (int t0 = 0; t0 < cnt; t0 += TB) {
        const int tcnt = min(TB, cnt - t0);
#pragma unroll
        for (int tt = 0; tt < TB; tt += U) {
          float4 v[U][SLOTS];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const bool on = tt + u < tcnt;
            const int sidx = on ? sh_src[gbase + t0 + tt + u] : 0;
            const float* rowp = P.wh + (int64_t)sidx * P.dp + gl * 4;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
              v[u][s] = (on && ok[s]) ? ldg4(rowp + s * G * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (tt + u < tcnt) {
#pragma unroll
              for (int s = 0; s < SLOTS; ++s) {
                if (ok[s]) {
                  float d = go[s].x * v[u][s].x;
                  d = fmaf(go[s].y, v[u][s].y, d);
                  d = fmaf(go[s].z, v[u][s].z, d);
                  d = fmaf(go[s].w, v[u][s].w, d);
                  part[(tt + u) * pstride + s * G + gl] = d;
                }
              }
            }
          }
        }
        __syncwarp(gmask);
        // transpose-reduce: (edge, head) pair q sums the chunks of that head
        for (int q = gl; q < tcnt * nh; q += G) {
          const int t = q / nh, h = q - t * nh;
          const float* pp = part + t * pstride + h * P.chunks_per_head;
          float d = 0.f;
          for (int c = 0; c < P.chunks_per_head; ++c) d += pp[c];
          sh_da[(gbase + t0 + t) * kMaxHeads + h] = d;
        }
        __syncwarp(gmask);
      }
      if (valid) {
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h) {
          if (h < nh) {
            dal[h] = fmaf(msk[h], sh_da[tid * kMaxHeads + h], ga[h]);
            ssum[h] = fmaf(alpha[h], dal[h], ssum[h]);
          }
        }
        if (!single) {   // stage {d_alpha, alpha} in the record slot; finalised in pass 2
          float* r = P.rec + (int64_t)e * 2 * nh;
#pragma unroll
          for (int h = 0; h < kMaxHeads; ++h)
            if (h < nh) { r[h] = dal[h]; r[nh + h] = alpha[h]; }
        }
      }
      __syncwarp(gmask);
    }
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h)
      if (h < nh) ssum[h] = group_sum<G>(ssum[h], gmask);

    // ---- pass 2: g = slope * alpha * (d_alpha - S);  record {g, m*alpha};  ds_tgt = sum g
    float gsum[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) gsum[h] = 0.f;
    for (int base = start; base < end; base += G) {
      const int e = base + gl;
      if (e < end) {
        float* r = P.rec + (int64_t)e * 2 * nh;
        if (!single) {
#pragma unroll
          for (int h = 0; h < kMaxHeads; ++h)
            if (h < nh) { dal[h] = r[h]; alpha[h] = r[nh + h]; }
          if (P.dropout_p > 0.f) dropout_scales(P.seed, P.offset, (uint32_t)__ldg(P.eid + e), nh, P.dropout_p, msk);
        }
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h) {
          if (h < nh) {
            // LeakyReLU'(l - M) = 0.01 everywhere: l - M <= 0, and torch uses the slope at exactly 0
            const float g = P.const_attention ? 0.f : kLeakySlope * alpha[h] * (dal[h] - ssum[h]);
            r[h] = g;
            r[nh + h] = msk[h] * alpha[h];
            gsum[h] += g;
            gam += (double)g;
          }
        }
      }
    }
    if (!P.const_attention) {
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h)
        if (h < nh) gsum[h] = group_sum<G>(gsum[h], gmask);
      if (gl == 0) {
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h)
          if (h < nh) P.ds_tgt[row * nh + h] = gsum[h];
      }
    }
  }

  // ---- fixed-order block reduction of the g partial sums
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gam += __shfl_xor_sync(0xffffffffu, gam, o);
  if (lane == 0) sh_gamma[tid >> 5] = gam;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kEdgeThreads / 32; ++w) t += sh_gamma[w];
    P.partials[blockIdx.x] = t;
    if (blockIdx.x == 0) P.header->n_partials = (int)gridDim.x;
  }
}

// Gamma = sum of the dst-pass partials (fixed order); corr = Gamma / |T|.
__global__ void __launch_bounds__(1024)
gamma_finalize_kernel(BwdHeader* header, const double* __restrict__ partials, const unsigned long long* __restrict__ tie_total) {
  __shared__ double sh[1024];
  const int n = header->n_partials;
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) t += partials[i];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    header->gamma = sh[0];
    unsigned long long ties = tie_total ? *tie_total : 0ull;
    header->corr = ties ? (float)(sh[0] / (double)ties) : 0.f;
  }
}

struct EdgeBwdSrcParams {
  const int32_t* rowptr_t; const int32_t* col_t; const int32_t* pos_t; int64_t n;
  int nh; int dp; int chunks; int chunks_per_head;
  const float* rec; const float* go; const float* a_src; const float* a_tgt; int const_attention;
  const int32_t* tie_dst; const int32_t* tie_src; const BwdHeader* header; const float* corr_override;
  int64_t tgt_lo; int64_t tgt_hi;   // rows that this call owns as TARGETS (ds_tgt / tie_dst are indexed row - tgt_lo)
  float* ds_src; float* ds_tgt; float* d_wh;
};

template <int G, int SLOTS>
__global__ void __launch_bounds__(kEdgeThreads)
edge_bwd_src_kernel(const EdgeBwdSrcParams P) {
  constexpr int U = SLOTS >= 4 ? 2 : (SLOTS >= 2 ? 4 : 8);
  __shared__ int sh_dst[kEdgeThreads];
  __shared__ float sh_w[kEdgeThreads * kMaxHeads];
  const int tid = threadIdx.x, lane = tid & 31, gl = tid & (G - 1), gbase = tid - gl;
  const unsigned gmask = group_mask<G>(lane);
  const int64_t row = (int64_t)blockIdx.x * (kEdgeThreads / G) + tid / G;
  if (row >= P.n) return;
  const int nh = P.nh;
  int head[SLOTS];
  bool ok[SLOTS];
  float4 acc[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    int c = s * G + gl;
    ok[s] = c < P.chunks;
    head[s] = ok[s] ? c / P.chunks_per_head : 0;
    acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int start = __ldg(P.rowptr_t + row), end = __ldg(P.rowptr_t + row + 1);
  float gsum[kMaxHeads];
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) gsum[h] = 0.f;

  for (int base = start; base < end; base += G) {
    const int e = base + gl;
    if (e < end) {
      sh_dst[tid] = __ldg(P.col_t + e);
      const float* r = P.rec + (int64_t)__ldg(P.pos_t + e) * 2 * nh;
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h) {
        float g = 0.f, w = 0.f;
        if (h < nh) { g = __ldg(r + h); w = __ldg(r + nh + h); }
        gsum[h] += g;
        sh_w[tid * kMaxHeads + h] = w;
      }
    }
    __syncwarp(gmask);
    const int cnt = min(G, end - base);
    for (int t = 0; t < cnt; t += U) {
      float4 v[U][SLOTS];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool on = t + u < cnt;
        const int didx = on ? sh_dst[gbase + t + u] : 0;
        const float* rowp = P.go + (int64_t)didx * P.dp + gl * 4;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
          v[u][s] = (on && ok[s]) ? ldg4(rowp + s * G * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (t + u < cnt) {
#pragma unroll
          for (int s = 0; s < SLOTS; ++s) {
            const float w = sh_w[(gbase + t + u) * kMaxHeads + head[s]];
            acc[s].x = fmaf(w, v[u][s].x, acc[s].x);
            acc[s].y = fmaf(w, v[u][s].y, acc[s].y);
            acc[s].z = fmaf(w, v[u][s].z, acc[s].z);
            acc[s].w = fmaf(w, v[u][s].w, acc[s].w);
          }
        }
      }
    }
    __syncwarp(gmask);
  }

  if (!P.const_attention) {
    // ds_src = sum g - |T_src|*Gamma/|T|;  ds_tgt -= |T_dst|*Gamma/|T|   (gradient through max(), section 9.2)
    const float corr = P.corr_override ? __ldg(P.corr_override) : P.header->corr;
    const bool own_tgt = row >= P.tgt_lo && row < P.tgt_hi;   // single GPU: always; partitioned: owner rank only
    const int64_t trow = row - P.tgt_lo;
    float dss[kMaxHeads], dst_[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) {
      dss[h] = 0.f; dst_[h] = 0.f;
      if (h < nh) {
        float g = group_sum<G>(gsum[h], gmask);
        int ts = P.tie_src ? __ldg(P.tie_src + row * nh + h) : 0;
        dss[h] = ts ? g - (float)ts * corr : g;
        if (own_tgt) {
          int td = P.tie_dst ? __ldg(P.tie_dst + trow * nh + h) : 0;
          float t = P.ds_tgt[trow * nh + h];
          dst_[h] = td ? t - (float)td * corr : t;
        }
      }
    }
    __syncwarp(gmask);   // every lane has read ds_tgt before lane 0 overwrites it
    if (gl == 0) {
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h)
        if (h < nh) {
          P.ds_src[row * nh + h] = dss[h];
          if (own_tgt) P.ds_tgt[trow * nh + h] = dst_[h];
        }
    }
    // d_wh_total = d_wh + ds_src * A_src + ds_tgt * A_tgt
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      if (ok[s]) {
        const int c4 = (s * G + gl) * 4;
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h) {
          if (h < nh) {
            const float4 as = ldg4(P.a_src + (int64_t)h * P.dp + c4);
            const float4 at = ldg4(P.a_tgt + (int64_t)h * P.dp + c4);
            acc[s].x = fmaf(dss[h], as.x, fmaf(dst_[h], at.x, acc[s].x));
            acc[s].y = fmaf(dss[h], as.y, fmaf(dst_[h], at.y, acc[s].y));
            acc[s].z = fmaf(dss[h], as.z, fmaf(dst_[h], at.z, acc[s].z));
            acc[s].w = fmaf(dss[h], as.w, fmaf(dst_[h], at.w, acc[s].w));
          }
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < SLOTS; ++s)
    if (ok[s]) *reinterpret_cast<float4*>(P.d_wh + row * P.dp + (s * G + gl) * 4) = acc[s];
}

template <int G, int SLOTS>
static size_t dst_dyn_smem(int chunks) {
  constexpr int TB = (G < 8) ? G : (SLOTS >= 6 ? 4 : 8);
  return (size_t)(kEdgeThreads / G) * TB * (chunks + 1) * sizeof(float);
}

}  // namespace gat

extern "C" size_t gat_edge_bwd_workspace_bytes(int64_t n, int64_t n_edges, int nh) {
  (void)n_edges; (void)nh;
  return gat::kBwdHeaderBytes + (size_t)(n + 1) * sizeof(double);   // G=1 worst case: one partial per 256 rows, G=32: per 8 rows
}

extern "C" int gat_edge_bwd_dst(const int32_t* rowptr, const int32_t* col, const int32_t* eid, int64_t n,
                                const float* wh, int nh, int fp, const float* s_src, const float* s_tgt,
                                const float* gmax, const float* z, int const_attention,
                                float dropout_p, uint64_t seed, uint64_t offset,
                                const float* go_padded, const float* grad_alpha,
                                float* rec, float* ds_tgt, void* workspace, size_t workspace_bytes,
                                gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(nh >= 1 && nh <= kMaxHeads, "gat_edge_bwd_dst: num_heads %d not in [1, %d]", nh, kMaxHeads);
  GAT_CHECK_ARG(fp > 0 && fp % 4 == 0, "gat_edge_bwd_dst: padded head width %d must be a positive multiple of 4", fp);
  GAT_CHECK_ARG(const_attention || (s_src && s_tgt && gmax), "gat_edge_bwd_dst: score buffers missing");
  if (workspace == nullptr || workspace_bytes < gat_edge_bwd_workspace_bytes(n, 0, nh)) {
    set_error("gat_edge_bwd_dst: workspace too small");
    return GAT_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  GAT_CUDA(cudaMemsetAsync(workspace, 0, kBwdHeaderBytes, st));
  if (n == 0) return GAT_OK;
  EdgeBwdDstParams P;
  P.rowptr = rowptr; P.col = col; P.eid = eid; P.n = n; P.wh = wh; P.nh = nh; P.dp = nh * fp;
  P.chunks = nh * fp / 4; P.chunks_per_head = fp / 4;
  P.s_src = s_src; P.s_tgt = s_tgt; P.gmax = gmax; P.z = z; P.const_attention = const_attention;
  P.dropout_p = dropout_p; P.seed = seed; P.offset = offset; P.go = go_padded; P.grad_alpha = grad_alpha;
  P.rec = rec; P.ds_tgt = ds_tgt; P.header = (BwdHeader*)workspace;
  P.partials = (double*)((char*)workspace + kBwdHeaderBytes);
  GroupShape shape = pick_group(P.chunks);
  if (shape.slots < 0) {
    set_error("gat_edge_bwd_dst: row width %d floats exceeds the supported 1024", P.dp);
    return GAT_EUNSUPPORTED;
  }
#define LAUNCH(G_, S_)                                                                                   \
  edge_bwd_dst_kernel<G_, S_><<<(unsigned)((n + (kEdgeThreads / G_) - 1) / (kEdgeThreads / G_)), kEdgeThreads, \
                                dst_dyn_smem<G_, S_>(P.chunks), st>>>(P)
  GAT_DISPATCH_GROUP(shape, LAUNCH);
#undef LAUNCH
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_edge_bwd_gamma(void* workspace, size_t workspace_bytes, double* gamma_out, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(workspace && gamma_out && workspace_bytes >= kBwdHeaderBytes, "gat_edge_bwd_gamma: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  BwdHeader* header = (BwdHeader*)workspace;
  gamma_finalize_kernel<<<1, 1024, 0, st>>>(header, (const double*)((char*)workspace + kBwdHeaderBytes), nullptr);
  GAT_LAUNCH_CHECK();
  GAT_CUDA(cudaMemcpyAsync(gamma_out, &header->gamma, sizeof(double), cudaMemcpyDeviceToDevice, st));
  return GAT_OK;
}

extern "C" int gat_edge_bwd_src(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* pos_t, int64_t n,
                                int nh, int fp, const float* rec, const float* go_padded,
                                const float* a_src, const float* a_tgt, int const_attention,
                                const int32_t* tie_dst, const int32_t* tie_src, const unsigned long long* tie_total,
                                const float* corr_override, int64_t tgt_lo, int64_t tgt_hi,
                                float* ds_src, float* ds_tgt, float* d_wh,
                                void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(nh >= 1 && nh <= kMaxHeads, "gat_edge_bwd_src: num_heads %d not in [1, %d]", nh, kMaxHeads);
  GAT_CHECK_ARG(fp > 0 && fp % 4 == 0, "gat_edge_bwd_src: padded head width %d must be a positive multiple of 4", fp);
  GAT_CHECK_ARG(const_attention || (a_src && a_tgt && ds_src && ds_tgt), "gat_edge_bwd_src: attention buffers missing");
  if (workspace == nullptr || workspace_bytes < gat_edge_bwd_workspace_bytes(n, 0, nh)) {
    set_error("gat_edge_bwd_src: workspace too small");
    return GAT_EWORKSPACE;
  }
  if (n == 0) return GAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  BwdHeader* header = (BwdHeader*)workspace;
  if (!const_attention && corr_override == nullptr) {
    gamma_finalize_kernel<<<1, 1024, 0, st>>>(header, (const double*)((char*)workspace + kBwdHeaderBytes), tie_total);
    GAT_LAUNCH_CHECK();
  }
  EdgeBwdSrcParams P;
  P.rowptr_t = rowptr_t; P.col_t = col_t; P.pos_t = pos_t; P.n = n; P.nh = nh; P.dp = nh * fp;
  P.chunks = nh * fp / 4; P.chunks_per_head = fp / 4;
  P.rec = rec; P.go = go_padded; P.a_src = a_src; P.a_tgt = a_tgt; P.const_attention = const_attention;
  P.tie_dst = tie_dst; P.tie_src = tie_src; P.header = header; P.corr_override = corr_override;
  P.tgt_lo = tgt_lo; P.tgt_hi = tgt_hi;
  P.ds_src = ds_src; P.ds_tgt = ds_tgt; P.d_wh = d_wh;
  GroupShape shape = pick_group(P.chunks);
  if (shape.slots < 0) {
    set_error("gat_edge_bwd_src: row width %d floats exceeds the supported 1024", P.dp);
    return GAT_EUNSUPPORTED;
  }
#define LAUNCH(G_, S_)                                                                          \
  edge_bwd_src_kernel<G_, S_><<<(unsigned)((n + (kEdgeThreads / G_) - 1) / (kEdgeThreads / G_)), kEdgeThreads, 0, st>>>(P)
  GAT_DISPATCH_GROUP(shape, LAUNCH);
#undef LAUNCH
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}
